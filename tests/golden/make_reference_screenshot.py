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


# ---- the second (and last) screenshot: Doc/Images/VoxelizationHiRes.jpg ("Hi-resolution example (not default)",
# README.md:10; the grid size is not stated).  Besides the silhouette it shows what ONLY the reference's shader
# produces: a spray of stray voxels behind the base of the bunny's left ear -- radial rays whose closest hit passes the
# normal threshold although the voxel is outside (DXRVoxelizer.hlsl:132-140).  The speckles are faint (|rgb - clear|
# summed over the channels = 20..40 of 765) and a few pixels wide, so they are kept at full resolution: the fixture
# holds the bit-packed masks `deviation > 20` and `> 30` of the 1920 x 1080 client area (they compress to a few KB)
# and the same 480 x 270 RGB thumbnail as above.
src2 = "/root/reference/Doc/Images/VoxelizationHiRes.jpg"
im2 = np.asarray(Image.open(src2).convert("RGB"))
client2 = im2[47:1127, 1:1921]
assert client2.shape == (1080, 1920, 3)
clear = np.array([0, 51, 102], np.float32)                        # CLEAR_COLOR (SharedConst.h:8) as UNORM8
dev = np.abs(client2.astype(np.float32) - clear).sum(-1)
small2 = np.asarray(Image.fromarray(client2).resize((480, 270), Image.BILINEAR))
out2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_screenshot_bunny_hires.npz")
np.savez_compressed(out2, rgb=small2, mask20=np.packbits(dev > 20, axis=1), mask30=np.packbits(dev > 30, axis=1),
                    shape=np.array([1080, 1920]), source=np.array(src2), crop=np.array([47, 1127, 1, 1921]))
print(out2, os.path.getsize(out2))
