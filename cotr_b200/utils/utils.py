"""Boundary helpers the inference path and the demos use (reference: COTR/utils/utils.py).

Only what sits on the hot path or on `demo_single_pair.py`'s call sequence is provided: ImagePatch (utils.py:24),
fix_randomness (:57-62), float_image_resize (:69-83), has_nan (:95-103), the torch<->numpy image layout helpers
(:128-160), safe_load_weights (:164-193).  Plotting lives in visualize_corrs, which degrades to a cv2 canvas when
matplotlib is absent (visualisation is out of scope).
"""
import random
from collections import namedtuple

import numpy as np
import PIL.Image
import torch

# patch: pixel content (or None); x, y: upper-left corner in the source image; w, h: patch size;
# ow, oh: size of the source image
ImagePatch = namedtuple('ImagePatch', ['patch', 'x', 'y', 'w', 'h', 'ow', 'oh'])


def fix_randomness(seed=42):
    random.seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.manual_seed(seed)
    np.random.seed(seed)


def float_image_resize(img, shape, interp=PIL.Image.BILINEAR):
    """Resize a float image (H,W) or (H,W,C) to `shape` = (H',W') channel by channel with Pillow's mode-'F' filter."""
    arr = img[..., None] if img.ndim == 2 else img
    planes = []
    for c in range(arr.shape[2]):
        plane = np.array(PIL.Image.fromarray(arr[..., c]).resize(tuple(shape)[::-1], resample=interp))
        assert plane.shape[:2] == tuple(shape)
        planes.append(plane)
    out = np.stack(planes, axis=-1)
    return out[..., 0] if img.ndim == 2 else out


def is_nan(x):
    return x != x


def has_nan(x) -> bool:
    return False if x is None else bool(is_nan(x).any())


def print_notification(content_list, notification_type='NOTIFICATION'):
    print('---------------------- {0} ----------------------'.format(notification_type))
    print()
    for content in content_list:
        print(content)
    print()
    print('----------------------------------------------------')


def torch_img_to_np_img(torch_img):
    """CHW (or NCHW) torch image -> HWC (NHWC) numpy image."""
    assert isinstance(torch_img, torch.Tensor), f'cannot process data type: {type(torch_img)}'
    a = torch_img.detach().cpu().numpy()
    if a.ndim == 4 and a.shape[1] in (1, 3):
        return a.transpose(0, 2, 3, 1)
    if a.ndim == 3 and a.shape[0] in (1, 3):
        return a.transpose(1, 2, 0)
    if a.ndim == 2:
        return a
    raise ValueError('cannot process this image')


def np_img_to_torch_img(np_img):
    """HWC (or NHWC) numpy image -> CHW (NCHW) torch image."""
    assert isinstance(np_img, np.ndarray), f'cannot process data type: {type(np_img)}'
    if np_img.ndim == 4 and np_img.shape[3] in (1, 3):
        return torch.from_numpy(np_img.transpose(0, 3, 1, 2))
    if np_img.ndim == 3 and np_img.shape[2] in (1, 3):
        return torch.from_numpy(np_img.transpose(2, 0, 1))
    if np_img.ndim == 2:
        return torch.from_numpy(np_img)
    raise ValueError(f'cannot process this image with shape: {np_img.shape}')


def safe_load_weights(model, saved_weights):
    """Load a checkpoint's model_state_dict: strict, then without / with a 'module.' prefix, then shape-matched partial."""
    attempts = (
        lambda: saved_weights,
        lambda: {k.replace('module.', ''): v for k, v in saved_weights.items()},
        lambda: {'module.' + k: v for k, v in saved_weights.items()},
    )
    for make in attempts:
        try:
            model.load_state_dict(make())
            print('weights safely loaded')
            return
        except RuntimeError:
            continue
    try:
        own = model.state_dict()
        usable = {k: v for k, v in saved_weights.items() if k in own and own[k].shape == v.shape}
        assert len(usable) != 0
        own.update(usable)
        model.load_state_dict(own)
        missing = set(model.state_dict().keys()) - set(usable.keys())
        print_notification(['pretrained weights PARTIALLY loaded, following are missing:', str(missing)], 'WARNING')
    except Exception as e:
        print(f'pretrained weights loading failed {e}')
        exit()
    print('weights safely loaded')


def visualize_corrs(img1, img2, corrs, mask=None, out_path=None):
    """Draw correspondences on a side-by-side canvas.  Returns the canvas (and writes it if out_path is given)."""
    import cv2
    if mask is None:
        mask = np.ones(len(corrs)).astype(bool)
    h = max(img1.shape[0], img2.shape[0])
    canvas = np.zeros((h, img1.shape[1] + img2.shape[1], 3), dtype=np.uint8)
    canvas[:img1.shape[0], :img1.shape[1]] = img1[..., :3]
    canvas[:img2.shape[0], img1.shape[1]:] = img2[..., :3]
    for (xa, ya, xb, yb), ok in zip(np.asarray(corrs)[:, :4], mask):
        color = (0, 255, 0) if ok else (255, 0, 0)
        cv2.line(canvas, (int(xa), int(ya)), (int(xb) + img1.shape[1], int(yb)), color, 1, cv2.LINE_AA)
    if out_path is not None:
        cv2.imwrite(out_path, canvas[..., ::-1])
    return canvas
