"""Generates tests/golden/example_png.npz from the only rendered artefact the reference ships.

Run in the authoring container (needs /root/reference):  python tests/golden/make_golden.py
The fixture is the decoded RGB8 pixel array of /root/reference/images/example.png (640x480),
stored losslessly, so the golden tests can run where /root/reference does not exist.
"""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
src = "/root/reference/images/example.png"
px = np.asarray(Image.open(src).convert("RGB"), dtype=np.uint8)
assert px.shape == (480, 640, 3)
np.savez_compressed(os.path.join(HERE, "example_png.npz"), rgb=px, source=np.array(src))
print("wrote", os.path.join(HERE, "example_png.npz"), px.shape)
