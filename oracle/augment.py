"""Training-data augmentation of the reference (data/base_dataset.py:87-131 with
--preprocess random3Drotate_randomcrop_randomflip_addColorChannel_addBatchChannel, README training command) restated
on the CPU.  Test infrastructure — see oracle/__init__.py.  Groundwork for SURVEY.md §8(f) item 1: the reference
re-rotates EVERY z-slice of the whole volume with cv2 on every iteration (1.67 s of its 3.65 s iteration); the
restatement evaluates only the voxels of the requested crop and is bit-identical to the reference pipeline.

The arithmetic of the rotation lives in a third-party dependency (OpenCV 4.5.0, conda_environment/neuroclear_env.yml:159:
cv2.getRotationMatrix2D + cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0)); `warp_affine_u16` restates its published
algorithm (imgwarp.cpp: inverse map in 1/1024 fixed point, 1/32 sub-pixel weights from a float table, float
accumulation, round-half-even cast) and is pinned against the cv2 installed in the build container
(oracle/make_golden.py::golden_augment) and against the reference's own transform pipeline.
"""
from __future__ import annotations

import math

import numpy as np

AB_BITS, INTER_BITS = 10, 5
AB_SCALE, INTER_TAB_SIZE = 1 << AB_BITS, 1 << INTER_BITS
ROUND_DELTA = AB_SCALE // INTER_TAB_SIZE // 2


def rotation_matrix_2d(center, angle_deg, scale=1.0):
    """cv2.getRotationMatrix2D"""
    a = math.radians(angle_deg)
    alpha, beta = scale * math.cos(a), scale * math.sin(a)
    return np.array([[alpha, beta, (1 - alpha) * center[0] - beta * center[1]],
                     [-beta, alpha, beta * center[0] + (1 - alpha) * center[1]]], dtype=np.float64)


def rotate_plan(width, height, angle_deg):
    """rotate_image (base_dataset.py:306-372): the 2x3 forward matrix and the size of the enlarged canvas."""
    cx, cy = width / 2, height / 2
    rot = np.vstack([rotation_matrix_2d((cx, cy), angle_deg), [0, 0, 1]])
    r2 = rot[0:2, 0:2]
    w2, h2 = width * 0.5, height * 0.5
    corners = [np.array(c) @ r2 for c in ([-w2, h2], [w2, h2], [-w2, -h2], [w2, -h2])]
    xs, ys = [c[0] for c in corners], [c[1] for c in corners]
    right, left = max(x for x in xs if x > 0), min(x for x in xs if x < 0)
    top, bot = max(y for y in ys if y > 0), min(y for y in ys if y < 0)
    new_w, new_h = int(abs(right - left)), int(abs(top - bot))
    trans = np.array([[1, 0, int(new_w * 0.5 - w2)], [0, 1, int(new_h * 0.5 - h2)], [0, 0, 1]], dtype=np.float64)
    return (trans @ rot)[0:2, :], new_w, new_h


def largest_rotated_rect(w, h, angle):
    """base_dataset.py:375-408 (as written there, including its gamma expression)"""
    quadrant = int(math.floor(angle / (math.pi / 2))) & 3
    sign_alpha = angle if ((quadrant & 1) == 0) else math.pi - angle
    alpha = (sign_alpha % math.pi + math.pi) % math.pi
    bb_w = w * math.cos(alpha) + h * math.sin(alpha)
    bb_h = w * math.sin(alpha) + h * math.cos(alpha)
    gamma = math.atan2(bb_w, bb_w)
    delta = math.pi - alpha - gamma
    length = h if (w < h) else w
    d = length * math.cos(alpha)
    a = d * math.sin(alpha) / math.sin(delta)
    y = a * math.cos(gamma)
    x = y * math.tan(gamma)
    return bb_w - 2 * x, bb_h - 2 * y


def center_crop_window(canvas_w, canvas_h, width, height):
    """crop_around_center (base_dataset.py:411-431) -> (x1, x2, y1, y2)"""
    cx, cy = int(canvas_w * 0.5), int(canvas_h * 0.5)
    width, height = min(width, canvas_w), min(height, canvas_h)
    return int(cx - width * 0.5), int(cx + width * 0.5), int(cy - height * 0.5), int(cy + height * 0.5)


def invert_affine(m):
    """the inversion cv2.warpAffine applies when WARP_INVERSE_MAP is not set (imgwarp.cpp)"""
    m = np.array(m, dtype=np.float64)
    d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = a11, m[0, 1] * -d, m[1, 0] * -d, a22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2], m[1, 2] = b1, b2
    return m


def _bilinear_tab():
    t = np.arange(INTER_TAB_SIZE, dtype=np.float32) / np.float32(INTER_TAB_SIZE)
    c = np.stack([np.float32(1) - t, t], 1)                      # interpolateLinear: (1 - x, x) in float32
    return c                                                       # tab[fy][fx][i][j] = c[fy][i] * c[fx][j]


def warp_affine_u16(src_stack, fwd_matrix, xs, ys):
    """cv2.warpAffine(src, M, dsize, INTER_LINEAR) (BORDER_CONSTANT, value 0) for uint16 images, evaluated only at
    the destination pixels (xs[j], ys[i]).  src_stack: (..., H, W) uint16 (leading axes = independent images sharing
    the matrix).  Returns (..., len(ys), len(xs)) uint16."""
    src = np.asarray(src_stack)
    assert src.dtype == np.uint16
    h, w = src.shape[-2:]
    m = invert_affine(fwd_matrix)
    xs, ys = np.asarray(xs, dtype=np.int64), np.asarray(ys, dtype=np.int64)
    rint = lambda v: np.rint(v).astype(np.int64)                 # cvRound: round half to even
    adelta, bdelta = rint(m[0, 0] * xs * AB_SCALE), rint(m[1, 0] * xs * AB_SCALE)
    x0 = rint((m[0, 1] * ys + m[0, 2]) * AB_SCALE) + ROUND_DELTA
    y0 = rint((m[1, 1] * ys + m[1, 2]) * AB_SCALE) + ROUND_DELTA
    X = (x0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx, sy = np.clip(X >> INTER_BITS, -32768, 32767), np.clip(Y >> INTER_BITS, -32768, 32767)   # saturate_cast<short>
    fx, fy = X & (INTER_TAB_SIZE - 1), Y & (INTER_TAB_SIZE - 1)
    c = _bilinear_tab()
    w00, w01 = c[fy, 0] * c[fx, 0], c[fy, 0] * c[fx, 1]          # float32 products, as the table is built
    w10, w11 = c[fy, 1] * c[fx, 0], c[fy, 1] * c[fx, 1]

    def fetch(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[..., np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.float32)
        return np.where(ok, v, np.float32(0))

    acc = fetch(sy, sx) * w00
    acc = acc + fetch(sy, sx + 1) * w01
    acc = acc + fetch(sy + 1, sx) * w10
    acc = acc + fetch(sy + 1, sx + 1) * w11                        # left-to-right float32 sum, as in remapBilinear
    return np.clip(np.rint(acc), 0, 65535).astype(np.uint16)       # saturate_cast<ushort>(float)


def rotate_clean_window(height, width, angle_deg):
    """__rotate_clean (base_dataset.py:433-443): forward matrix and the window [y1:y2, x1:x2] of the rotated canvas
    the reference keeps."""
    m, new_w, new_h = rotate_plan(width, height, angle_deg)
    rw, rh = largest_rotated_rect(width, height, math.radians(angle_deg))
    x1, x2, y1, y2 = center_crop_window(new_w, new_h, rw, rh)
    # numpy slicing of the canvas clips at its edges
    return m, max(x1, 0), min(x2, new_w), max(y1, 0), min(y2, new_h)


def rotated_volume_shape(vol_shape, angle_deg):
    _, x1, x2, y1, y2 = rotate_clean_window(vol_shape[1], vol_shape[2], angle_deg)
    return vol_shape[0], max(y2 - y1, 0), max(x2 - x1, 0)


def augment_crop(vol_u16, angle_deg, crop_pos, crop_size, flip_axis=None):
    """get_transform(opt, params)(vol) for the README's --preprocess (rotate every slice by angle_3D, keep the largest
    axis-aligned rectangle, crop crop_size at crop_pos, normalise by 65535, flip one axis, add color + batch channel)
    evaluated ONLY on the crop: float32 (1, 1, cz, cy, cx)."""
    z, y, x = crop_pos
    cz, cy, cx = crop_size
    m, x1, x2, y1, y2 = rotate_clean_window(vol_u16.shape[1], vol_u16.shape[2], angle_deg)
    ys = np.arange(y1 + y, min(y1 + y + cy, y2))
    xs = np.arange(x1 + x, min(x1 + x + cx, x2))
    crop = warp_affine_u16(vol_u16[z:z + cz], m, xs, ys)
    out = (crop / (2 ** 16 * 1.0 - 1)).astype(float)              # __normalize (base_dataset.py:134-143)
    if flip_axis is not None:
        for ax in ([flip_axis] if np.isscalar(flip_axis) else flip_axis):
            out = np.flip(out, int(ax))
    return np.ascontiguousarray(out[None, None]).astype(np.float32)


def random_item(vol_u16, crop_size):
    """get_transform(opt)(vol) with params=None: the reference's random draws in its order — random.randint for the
    angle (__randomrotate_clean_3D_xy) and the crop position (__randomcrop), random.shuffle + np.random.uniform for
    the flips (__randomflip, base_dataset.py:279-289)."""
    import random
    angle = random.randint(0, 359)
    Z, Hr, Wr = rotated_volume_shape(vol_u16.shape, angle)
    cz, cy, cx = crop_size
    pos = (random.randint(0, Z - cz), random.randint(0, Hr - cy), random.randint(0, Wr - cx))
    axis_list = [0, 1, 2]
    random.shuffle(axis_list)
    flips = []
    for _ in range(3):
        if np.random.uniform(0, 1) < 0.5:
            flips.append(axis_list.pop())
    return augment_crop(vol_u16, angle, pos, crop_size, flips), (angle, pos, tuple(flips))
