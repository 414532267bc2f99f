"""CPU restatement of the reference's overlap-tile driver (TEST INFRASTRUCTURE -- only tests/, smoke() and bench.py's CPU
legs may import this).  Pinned against the unmodified reference (`oracle/ref_loader.load_reference_tiler`, live in the
build container: tests/test_oracle_golden.py) and against the fixture `tests/golden/tiler_prod.pt` minted from it by
`oracle/make_golden.py`.

Follows `hcat/segment.py:21-136` (`predict_segmentation_mask`), `hcat/utils.py:33-74` (`pad_image_with_reflections`) and
`hcat/utils.py:77-124` (`calculate_indexes`) step by step; `forward` is any callable tile -> logits (the reference model,
or `oracle.unet_oracle.unet_forward` in eval mode)."""
from __future__ import annotations

import math

import torch

EVAL_IM_SIZE = {"4": [128, 128, 6], "6": [300, 300, 6], "8": [300, 300, 10], "11": [350, 350, 15]}   # segment.py:48-51


def pad_image_with_reflections(image: torch.Tensor, pad_size=(30, 30, 6)) -> torch.Tensor:
    """utils.py:33-74: per dimension cat(first `pad` voxels reversed, image, last `pad` voxels reversed)."""
    if not isinstance(image, torch.Tensor):
        raise TypeError(f"Expected image to be of type torch.tensor not {type(image)}")
    for pad in pad_size:
        if pad % 2 != 0:
            raise ValueError("Padding must be divisible by 2")
    for dim, p in zip((2, 3, 4), pad_size):
        n = image.shape[dim]
        lo = image.narrow(dim, 0, p).flip(dim)            # image[p-1::-1]
        hi = image.narrow(dim, n - p, p).flip(dim)        # image[-1:-p-1:-1]
        image = torch.cat((lo, image, hi), dim=dim)
    return image


def calculate_indexes(pad_size: int, eval_image_size: int, image_shape: int, padded_image_shape: int):
    """utils.py:77-124."""
    if eval_image_size > image_shape:
        return [[0, image_shape]]
    ind_list = torch.arange(0, image_shape, eval_image_size)
    ind = []
    for i, z in enumerate(ind_list):
        if i == 0:
            continue
        z1 = int(ind_list[i - 1])
        z2 = int(z - 1) + (2 * pad_size)
        if z2 < padded_image_shape:
            ind.append([z1, z2])
        else:
            break
    if not ind:
        ind.append([0, eval_image_size + pad_size * 2])
        ind.append([padded_image_shape - (eval_image_size + pad_size * 2), padded_image_shape])
    else:
        ind.append([padded_image_shape - (eval_image_size + pad_size * 2), padded_image_shape - 1])
    return ind


def predict_segmentation_mask(forward, image: torch.Tensor, use_probability_map=False, mask_cell_prob_threshold=0.5,
                              cuda_mem=None):
    """segment.py:21-136 with `forward` in place of `unet(...)` (CPU).  Scrubs `image` in place like the reference."""
    if cuda_mem:
        PAD_SIZE = (128, 128, 10)
        EVAL = list(EVAL_IM_SIZE[str(int(math.floor(cuda_mem / 1e9)))])
    else:
        PAD_SIZE = [128, 128, 10]
        EVAL = [300, 300, 15]
    mask = torch.zeros((1, 1, image.shape[2], image.shape[3], image.shape[4]), dtype=torch.float)
    im_shape = image.shape
    if im_shape[4] < EVAL[2]:
        EVAL[2] = im_shape[4]
    image[torch.isnan(image)] = 0
    image[torch.isinf(image)] = 1
    image = pad_image_with_reflections(image, pad_size=PAD_SIZE)
    x_ind = calculate_indexes(PAD_SIZE[0], EVAL[0], im_shape[2], image.shape[2])
    y_ind = calculate_indexes(PAD_SIZE[1], EVAL[1], im_shape[3], image.shape[3])
    z_ind = calculate_indexes(PAD_SIZE[2], EVAL[2], im_shape[4], image.shape[4])
    skipped = 0
    with torch.no_grad():
        for z in z_ind:
            for x in x_ind:
                for y in y_ind:
                    sl = image[:, :, x[0]:x[1], y[0]:y[1], z[0]:z[1]].float()
                    if (sl == -1).all():
                        skipped += 1
                        continue
                    valid_out = forward(sl)
                    valid_out = valid_out[:, :, PAD_SIZE[0]:EVAL[0] + PAD_SIZE[0], PAD_SIZE[1]:EVAL[1] + PAD_SIZE[1],
                                          PAD_SIZE[2]:EVAL[2] + PAD_SIZE[2]].clone()
                    valid_out.mul_(-1).exp_().add_(1).pow_(-1)
                    if not use_probability_map:
                        valid_out.gt_(mask_cell_prob_threshold)
                        valid_out = valid_out.type(torch.uint8)
                        if mask.dtype != torch.uint8:
                            mask = mask.type(torch.uint8)
                    try:
                        mask[:, :, x[0]:x[0] + EVAL[0], y[0]:y[0] + EVAL[1], z[0]:z[0] + EVAL[2]] = valid_out
                    except (IndexError, RuntimeError):
                        raise RuntimeError(f"Amount of padding is not sufficient.\nvalid_out.shape: {valid_out.shape}\n"
                                           f"eval_image_size: {EVAL} ")
    return mask, skipped
