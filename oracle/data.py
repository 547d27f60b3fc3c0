"""CPU restatement of the reference's input pipeline (TEST INFRASTRUCTURE, see oracle/__init__.py).

    main.py:71-75      X = X.astype(np.float); X -= mean(X_train, axis=0); X /= 128      (float64; fp32 at the feed_dict)
    trainer.py:24-28   tf.image.random_flip_left_right; pad_to_bounding_box(image, 4, 4, 40, 40); tf.random_crop([32, 32, 3])

The reference's random draws (TF's unseeded ops) cannot be replayed, so the parity tests feed both sides explicit
(flip, oy, ox) triples, or the Philox draw defined in include/lbt.h (lbt_augment_batch).
"""
import numpy as np

from . import philox as P


def normalise(X_u8, mean64):
    """main.py:66-75: float64 arithmetic, one rounding to fp32 when fed to the float32 placeholder."""
    return ((X_u8.astype(np.float64) - mean64) / 128).astype(np.float32)


def preprocess_image(img, flip, oy, ox, pad=4):
    """trainer.py:24-28 for one HWC image with the random choices made explicit."""
    H, W, C = img.shape
    if flip:
        img = img[:, ::-1, :]                                   # random_flip_left_right: the width axis
    padded = np.zeros((H + 2 * pad, W + 2 * pad, C), dtype=img.dtype)
    padded[pad:pad + H, pad:pad + W] = img                      # pad_to_bounding_box(offset 4, 4)
    return padded[oy:oy + H, ox:ox + W]                         # random_crop


def philox_params(B, seed, offset, pad=4):
    """(flip, oy, ox) per sample as lbt_augment_batch draws them: counter (b, 0, offset), key seed."""
    seed, offset = int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1)
    b = np.arange(B, dtype=np.uint64)
    r0, r1, r2, _ = P.philox4x32_10(b, np.zeros(B, dtype=np.uint64), np.full(B, offset & 0xFFFFFFFF, dtype=np.uint64),
                                    np.full(B, offset >> 32, dtype=np.uint64), seed & 0xFFFFFFFF, seed >> 32)
    m = 2 * pad + 1
    return np.stack([(r0 & 1).astype(np.int32), (r1 % m).astype(np.int32), (r2 % m).astype(np.int32)], axis=1)


def batch(X_u8, y, mean64, index, params, pad=4, do_flip=True):
    """The batch lbt_augment_batch produces: NHWC fp32 images, labels."""
    Xn = normalise(X_u8[index], mean64)
    out = np.stack([preprocess_image(Xn[i], bool(params[i, 0]) and do_flip, int(params[i, 1]), int(params[i, 2]), pad)
                    for i in range(len(index))])
    return out, y[index]
