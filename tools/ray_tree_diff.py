"""Find where the GPU ray tree diverges from the oracle's: follow reflection / transmission chains of chosen pixels
with rtx_trace_probe on both back ends (ray geometry recomputed in float32 numpy, no contraction)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager
from oracle.oracle import OracleRenderer
from oracle import oracle
F = np.float32
name = sys.argv[1] if len(sys.argv) > 1 else "room_spheres"
W, H, S = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
fs, cam, cfg = abi.load_fixture(name, samples=S, monte_carlo=0)
cam = abi.resize_camera(cam, W, H)
g = RendererManager(W, H, fs); c = OracleRenderer(fs)
a = g.start(cam, cfg); sa = (a.stats.rays_closest, a.stats.rays_shadow); ia = a.image.copy()
b = g.start(cam, cfg); sb = (b.stats.rays_closest, b.stats.rays_shadow)
r = c.render(cam, cfg); sr = (r.stats.rays_closest, r.stats.rays_shadow)
print("gpu run1", sa, "gpu run2", sb, "cpu", sr)
d = np.abs(ia.astype(int) - r.image.astype(int)).max(-1)
ys, xs = np.nonzero(d > 1)
print("pixels >1 LSB:", len(ys), list(zip(xs[:10], ys[:10])), "ids equal", (a.objects == r.objects).all())

def dot(a, b): return F(F(F(a[0]*b[0]) + F(a[1]*b[1])) + F(a[2]*b[2]))
def norm(v): n = F(np.sqrt(dot(v, v))); return np.array([F(v[0]/n), F(v[1]/n), F(v[2]/n)], dtype=F)
mats = fs.materials
def follow(o, dd, depth, path, out):
    hg = g.trace([o], [dd], depth=depth)[0]; hc = c.trace([o], [dd], depth=depth)[0]
    same = hg.tobytes() == hc.tobytes()
    if not same:
        out.append((path, depth, o, dd, hg, hc)); return
    if hc["t"] < 0 or depth > cfg.max_recursion: return
    it = fs.items[int(hc["item_index"])]; m = mats[it.material]
    n = hc["normal"].astype(F); t = F(hc["t"])
    p = np.array([F(o[k] + F(dd[k]*t)) for k in range(3)], dtype=F)
    if m.reflectivity > 0:
        ro = np.array([F(p[k] + F(n[k]*F(0.001))) for k in range(3)], dtype=F)
        s2 = F(F(2.0) * dot(dd, n)); rd = np.array([F(dd[k] - F(n[k]*s2)) for k in range(3)], dtype=F)
        follow(ro, norm(rd), depth + 1, path * 2, out)
    alpha = F(m.alpha)   # (no textures with alpha in these scenes)
    if alpha < 1:
        idn = dot(dd, n); refn = n.copy(); eta_t, eta_i = F(m.refraction_index), F(1.0)
        if idn < 0: idn = F(-idn)
        else: refn = -n; eta_t, eta_i = F(1.0), F(m.refraction_index)
        eta = F(eta_i / eta_t); k = F(F(1.0) - F(F(eta*eta) * F(F(1.0) - F(idn*idn))))
        if k >= 0:
            to = np.array([F(p[j] + F(refn[j]*F(-0.001))) for j in range(3)], dtype=F)
            sq = F(np.sqrt(k))
            td = np.array([F(F(F(dd[j] + F(refn[j]*idn)) * eta) - F(refn[j]*sq)) for j in range(3)], dtype=F)
            follow(to, norm(td), depth + 1, path * 2 + 1, out)

cell, table = c.sample_table(S)
found = 0
for x, y in list(zip(xs, ys))[:40]:
    for s in range(S):
        o, dd = oracle.gen_ray(cam, cfg, int(x), int(y), int(table[s][0]), int(table[s][1]), cell)
        out = []
        follow(o.astype(F), norm(dd.astype(F)), 1, 1, out)
        for (path, depth, oo, d2, hg, hc) in out[:2]:
            found += 1
            print("pixel", x, y, "sample", s, "path", bin(path), "depth", depth, "\n   o", oo.tolist(), "d", d2.tolist(), "\n   gpu", hg, "\n   cpu", hc)
    if found > 6: break
print("mismatching probe hits found:", found)
