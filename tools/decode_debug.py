"""Bring-up helper for the decode fast path: runs with and without the CUDA graph under hard timeouts."""
import os
import subprocess
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(graph):
    os.environ["COGVIEW_B200_CUDA_GRAPH"] = graph
    import torch
    from oracle import recipes
    from cogview_b200.model import GPT2Model
    CFG = recipes.CONFIG1
    m = GPT2Model(num_layers=2, vocab_size=CFG["vocab_size"], hidden_size=256, num_attention_heads=4,
                  embedding_dropout_prob=0., attention_dropout_prob=0., output_dropout_prob=0.,
                  max_sequence_length=128, max_memory_length=128, checkpoint_activations=False)
    m.load_state_dict(recipes.gpt2_state_dict(**CFG))
    m = m.cuda().bfloat16().eval()
    toks = recipes.text_image_tokens(2, 64, 65, seed=0)[:, :64].cuda()
    pos = torch.arange(64, device="cuda").unsqueeze(0).expand(2, -1).contiguous()
    with torch.no_grad():
        lg, *mems = m(toks, pos, torch.tril(torch.ones((1, 1, 64, 64), device="cuda")), None, None, 0)
        torch.cuda.synchronize()
        print("prefill ok", lg.shape, mems[0].shape, flush=True)
        for t in range(64, 72):
            t0 = time.time()
            nxt = lg[:, -1, :8192].float().argmax(-1)
            lg, *mems = m(nxt.unsqueeze(1), torch.full((2, 1), t, dtype=torch.long, device="cuda"), 0, None, None, 0,
                          *mems)
            torch.cuda.synchronize()
            print("step", t, "ok %.3fs" % (time.time() - t0), mems[0].shape, nxt.tolist(), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for g in ("0", "1"):
            print("=== CUDA graph", g, flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, g], capture_output=True, text=True, timeout=120)
                print(r.stdout[-3000:], r.stderr[-3000:], flush=True)
            except subprocess.TimeoutExpired as e:
                print("TIMEOUT", (e.stdout or b"").decode()[-3000:], (e.stderr or b"").decode()[-2000:], flush=True)
