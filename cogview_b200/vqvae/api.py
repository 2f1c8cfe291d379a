"""Production tokenizer API — mirror of the reference's vqvae/api.py:12-44 (`new_model`, `img2code`, `code2img`)."""
import math

import torch

from .vqvae_zc import VQVAE

IMG_STD = (0.30379, 0.32279, 0.32800)
IMG_MEAN = (0.79093, 0.76271, 0.75340)


def new_model():
    """A VQVAE with the hyper-parameters of the released tokenizer (for torch.load / load_state_dict)."""
    return VQVAE(channel=512, n_res_block=0, n_res_channel=32, embed_dim=256, n_embed=8192, stride=6)


def img2code(model, img):
    """img: [b, c, h, w] -> codes [b, h/8 * w/8] (int64)."""
    with torch.no_grad():
        _, _, id_t1 = model.encode(img)
    return id_t1.view(img.shape[0], -1)


def code2img(model, code):
    """code: [b, h, w] or [b, h*w] LongTensor -> de-normalised image [b, 3, 8h, 8w]."""
    if len(code.shape) == 2:
        s = int(math.sqrt(len(code.view(-1))) + 1e-5)
        code = code.view(code.shape[0], s, s)
    with torch.no_grad():
        scale = torch.tensor(IMG_STD, device=code.device)
        shift = torch.tensor(IMG_MEAN, device=code.device)
        out = model.decode_code(code, scale, shift)     # the de-normalisation is fused into the last kernel
    return out
