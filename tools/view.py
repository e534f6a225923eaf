#!/usr/bin/env python3
"""Debug helper: XYZ float image (H,W,4) .npy -> PNG using the reference's display transform
(shader.frag:31-93: Bradford E->D65, XYZ->linear sRGB, tonemap 3 = exp(-0.25/x), sRGB companding)."""
import sys
import numpy as np
from PIL import Image


def xyz_to_srgb8(img, tonemap=3):
    xyz = img[..., :3].astype(np.float64)
    M_E_D65 = np.array([[0.9531874, -0.0265906, 0.0238731], [-0.0382467, 1.0288406, 0.0094060], [0.0026068, -0.0030332, 1.0892565]])
    M_RGB = np.array([[3.2404542, -1.5371385, -0.4985314], [-0.9692660, 1.8760108, 0.0415560], [0.0556434, -0.2040259, 1.0572252]])
    rgb = np.maximum(xyz @ M_E_D65.T @ M_RGB.T, 0.0)
    if tonemap == 3:
        with np.errstate(divide='ignore'):
            rgb = np.exp(-0.25 / np.maximum(rgb, 1e-12))
    elif tonemap == 1:
        rgb = rgb / (1 + rgb)
    rgb = np.clip(rgb, 0, 1)
    srgb = np.where(rgb <= 0.0031308, 12.92 * rgb, 1.055 * np.power(rgb, 1 / 2.4) - 0.055)
    return (np.clip(srgb, 0, 1) * 255).astype(np.uint8)


if __name__ == '__main__':
    img = np.load(sys.argv[1])
    Image.fromarray(xyz_to_srgb8(img)).save(sys.argv[2])
