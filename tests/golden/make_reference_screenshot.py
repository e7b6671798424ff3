#!/usr/bin/env python3
"""Fixture generator (run in the build container, where /root/reference exists): the client area of the reference's
own screenshot Doc/Images/SolidVoxelization.jpg (bunny, GRID_SIZE 64, 1920 x 1080 client area inside a 1922 x 1128
window capture), downsampled to 480 x 270 RGB.  It is the ONLY artefact of the voxel stage the reference tree holds.
Writes tests/golden/reference_screenshot_bunny64.npz."""
import os
import numpy as np
from PIL import Image

src = "/root/reference/Doc/Images/SolidVoxelization.jpg"
im = np.asarray(Image.open(src).convert("RGB"))
client = im[47:1127, 1:1921]                      # below the title bar, inside the 1-pixel window border
assert client.shape == (1080, 1920, 3)
small = np.asarray(Image.fromarray(client).resize((480, 270), Image.BILINEAR))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_screenshot_bunny64.npz")
np.savez_compressed(out, rgb=small, source=np.array(src), crop=np.array([47, 1127, 1, 1921]))
print(out, small.shape, os.path.getsize(out))
