"""Downsample three of the author's renderings (real outputs of the reference) into small fixtures: config 2
(floor + monkey), the room of glass / mirror spheres (room-no-textures.json + spheres.json) and kbert in the textured room
(kbert_in_room.json).
    python tests/golden/make_ref_render.py [/root/reference]"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = os.path.join(ref, "data", "renderings", "output_2022-5-16_20-47-31_00000000.png")
im = np.asarray(Image.open(src).convert("RGB")).astype(np.float32)
assert im.shape == (720, 1280, 3)
small = im.reshape(180, 4, 320, 4, 3).mean(axis=(1, 3))
Image.fromarray(np.clip(small + 0.5, 0, 255).astype(np.uint8)).save(os.path.join(HERE, "ref_render_c2_320x180.png"))
print("wrote ref_render_c2_320x180.png")
# ... and unscaled (174 KB): the GPU's own 32-spp Monte-Carlo frame is compared with it pixel for pixel (tests/test_gpu_round2.py)
import shutil
shutil.copyfile(src, os.path.join(HERE, "ref_render_c2_1280x720.png"))
print("wrote ref_render_c2_1280x720.png")

src = os.path.join(ref, "data", "renderings", "output_2022-5-16_21-24-33_00000000.png")
im = np.asarray(Image.open(src).convert("RGB")).astype(np.float32)
assert im.shape == (720, 1280, 3)
small = im.reshape(180, 4, 320, 4, 3).mean(axis=(1, 3))
Image.fromarray(np.clip(small + 0.5, 0, 255).astype(np.uint8)).save(os.path.join(HERE, "ref_render_room_spheres_320x180.png"))
print("wrote ref_render_room_spheres_320x180.png")

src = os.path.join(ref, "data", "renderings", "output_2022-5-16_15-41-8_00000000.png")
im = np.asarray(Image.open(src).convert("RGB")).astype(np.float32)
assert im.shape == (720, 1280, 3)
small = im.reshape(180, 4, 320, 4, 3).mean(axis=(1, 3))
Image.fromarray(np.clip(small + 0.5, 0, 255).astype(np.uint8)).save(os.path.join(HERE, "ref_render_kbert_in_room_320x180.png"))
print("wrote ref_render_kbert_in_room_320x180.png")
