"""BASELINE config 2 through the public API: MetaModel.generate_tokens, batch 1, 336 px image + 256-token
prompt, 128 generated tokens on one B200 -- greedy and with the reference's eval sampling settings
(temperature 0.1, top-p 0.75, M/eval_affordance_v2.py:46-49).  Wall clock around the whole call
(prefill + 128 decode steps + host loop), synchronised on both sides."""
import json, sys, time, torch
sys.path.insert(0, ".")
import bench
from a3vlm_b200.model import MetaModel
cfg = bench.MODELS["7b"]
model = MetaModel("llama_b200", {**cfg, "image_size": 336}, None, with_visual=True, max_seq_len=2048, max_batch_size=8)
model.llma.reset_parameters(seed=2)
g = torch.Generator().manual_seed(7)
res = {}
for B in (1, 8):
    prompts = torch.randint(3, 32000, (B, 256), generator=g); prompts[:, 0] = 1
    img = torch.randn(B, 3, 336, 336, generator=g).cuda()
    plist = prompts.tolist()
    for name, kw in (("greedy", dict(temperature=0.0)), ("sampled_t0.1_p0.75", dict(temperature=0.1, top_p=0.75))):
        best = None
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out, _, _ = model.generate_tokens(plist, img, max_gen_len=128, eos_id=-1, **kw)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        res[f"B{B}_{name}"] = {"generate_s": best, "tokens": sum(len(o) for o in out), "tokens_per_s": sum(len(o) for o in out) / best}
print(json.dumps(res))
