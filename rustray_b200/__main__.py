"""Stand-in for the reference CLI (reference src/main.rs:16-113, src/run.rs `cmd` path): same flags, same output file
naming, the frame rendered by librtx_b200.so on GPU 0.

    python -m rustray_b200 scene/floor.json scene/monkey.json cmd no-animation 1280x720 samples=32 monte_carlo=1

Flags: `cmd` (always headless here), `no-animation`, `monte_carlo=0|1|true`, `samples=N`, `WxH`, `start=1`, any number of
*.json / *.gltf / *.glb / *.obj scene files (loaded in order into ONE scene; a scene file's "config" block overrides the
CLI values, exactly like the reference — SURVEY.md fact 6).  Extra: `root=DIR` (asset root, default cwd), `out=DIR`.
"""
import datetime
import os
import re
import sys
import time


def main(argv):
    from . import abi
    from .animation import Animation
    from .renderer import RendererManager
    from .scene_loader import load_scene
    width, height = 800, 600                      # run.rs:34
    scenes, animation, monte_carlo, samples, root, out = [], True, None, None, ".", os.path.join("data", "output")
    for arg in argv:
        if arg == "cmd":
            pass
        elif arg == "no-animation":
            animation = False
        elif arg.startswith("monte_carlo="):
            monte_carlo = arg.split("=")[1] in ("1", "true")
        elif arg.endswith((".json", ".gltf", ".glb", ".obj")):
            scenes.append(arg)
        elif re.match(r"^\d+x\d+$", arg):
            width, height = (int(v) for v in arg.split("x"))
        elif arg.startswith("samples="):
            samples = int(arg.split("=")[1])
        elif arg.startswith("root="):
            root = arg.split("=", 1)[1]
        elif arg.startswith("out="):
            out = arg.split("=", 1)[1]
    if not scenes:
        print(__doc__)
        return 2
    sc = load_scene(scenes, width, height, asset_root=root, samples=samples, monte_carlo=monte_carlo)
    fs = abi.FlatScene.from_scene(sc)
    # under torchrun (one process per GPU) the frames of an animation are spread over the ranks, rank 0 writes the files
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = int(os.environ.get("LOCAL_RANK", "0"))
    rm = RendererManager(width, height, fs, device=device)
    cam, cfg = abi.make_camera(sc.cam), abi.make_config(sc.config, mc_seed=int(time.time()) & 0x7fffffff)
    anim = Animation(sc.animation)
    n_frames = anim.frames_to_render() if (animation and anim.has_animation()) else 1
    os.makedirs(out, exist_ok=True)
    from PIL import Image

    def render_frame(frame):
        if anim.has_animation():
            ups = anim.updates_for_frame(sc.items, frame)           # Scene::apply_frame (scene.rs:1695-1713)
            if ups:
                rm.update_items(ups)
        f = rm.start(cam, cfg)
        s = f.stats
        print("frame %d rendered ✅ (rendering time: %.3fs, %d closest + %d shadow rays, %.0f Mrays/s)" % (
            frame, s.device_ms / 1e3, s.rays_closest, s.rays_shadow, (s.rays_closest + s.rays_shadow) / max(s.device_ms, 1e-6) / 1e3))
        return f

    def save(frame, f):
        now = datetime.datetime.now()                                # run.rs:565-576
        name = "output_%d-%d-%d_%d-%d-%d_%08d.png" % (now.year, now.month, now.day, now.hour, now.minute, now.second, frame)
        Image.fromarray(f.image).save(os.path.join(out, name))

    if world > 1:
        import torch
        import torch.distributed as dist
        from .distributed import render_animation_frame_parallel
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        render_animation_frame_parallel(render_frame, n_frames, width, height, rank, world, on_frame=save, device=torch.device("cuda", device))
        dist.destroy_process_group()
    else:
        for frame in range(n_frames):
            save(frame, render_frame(frame))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
