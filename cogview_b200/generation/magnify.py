"""Super-resolution driver (generation/magnify.py:22-43): a 32 x 32 code image is magnified to 64 x 64 by nine
overlapping windows; each window conditions on the text, its 16 x 16 source patch and the already generated rows
of its 32-column strip, and fills the rest through `filling_sequence`."""
import math

import torch

from .sampling import filling_sequence

# (window row, window column, number of 32-code lines generated or re-read in that window)
WINDOWS = [(0, 0, 18), (0, 1, 30), (0, 2, 30), (1, 1, 30), (1, 0, 30), (1, 2, 30), (2, 0, 32), (2, 1, 32), (2, 2, 32)]


def magnify(model, tokenizer, tokens_list, text_token_list, args, fill=filling_sequence):
    """tokens_list: [1024] codes of the 32 x 32 image; text_token_list: 1-D text prefix.  Returns [1, 4096] codes."""
    s = int(math.sqrt(len(tokens_list) + 1e-6))
    assert s == 32
    code = tokens_list.view(s, s)
    midfix = torch.tensor([tokenizer['[EOI1]'], tokenizer['[ROI2]'], tokenizer['[POS0]'], tokenizer['[BASE]'],
                           tokenizer['[BOI2]']], device=code.device)
    magnified = code.new_zeros((s * 2, s * 2), dtype=torch.long) - 1
    for i, j, line in WINDOWS:
        code_part = code[8 * i: 8 * (i + 2), 8 * j: 8 * (j + 2)].reshape(-1)
        known = magnified[16 * i: 16 * i + line, 16 * j: 16 * (j + 2)].reshape(-1)      # -1 = still to generate
        context = torch.cat([text_token_list, code_part, midfix], dim=0)
        seq = torch.cat([context, known], dim=0)
        done = fill(model, seq, args, invalid_slices=[slice(tokenizer.img_tokenizer.num_tokens, None)])
        magnified[16 * i: 16 * i + line, 16 * j: 16 * (j + 2)] = done[0, len(context):].view(line, 32)
    return magnified.view(1, s * s * 4)
