"""BASELINE config 5 (single GPU part): inference throughput sweep, batch x image size x decode length,
Llama-2-7B + ViT-L/14 random init, 256-token prompts, greedy.  Also checks that the batch-16
(linear_small_m) and batch-64 (tcgen05 GEMM) decode paths generate the same ids as batch 1 for a
replicated prompt."""
import json, sys, time, torch
sys.path.insert(0, ".")
import bench
from a3vlm_b200.model import MetaModel
cfg = bench.MODELS["7b"]
dev = torch.device("cuda", 0)
combos = {224: [(1, 128), (16, 128)], 336: [(1, 32), (1, 128), (1, 512), (4, 128), (16, 128), (64, 128)], 448: [(1, 128), (16, 128)]}
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    combos = {336: [(1, 128), (2, 128), (4, 128), (8, 128), (16, 128), (64, 128)]}
FUSED_MAX_B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rows = []
for S, lst in combos.items():
    model = MetaModel("llama_b200", {**cfg, "image_size": S}, None, with_visual=True, max_seq_len=2048, max_batch_size=64)
    model.llma.reset_parameters(seed=2)
    eng = model.llma.engine
    eng.decode_fused_max_b = FUSED_MAX_B
    g = torch.Generator().manual_seed(7)
    base_prompt = torch.randint(3, 32000, (1, 256), generator=g); base_prompt[:, 0] = 1
    base_img = torch.randn(1, 3, S, S, generator=g)
    ref_ids = None
    for B, G in lst:
        prompts = base_prompt.repeat(B, 1).to(dev)
        img = base_img.repeat(B, 1, 1, 1).to(dev)
        eng.destroy_kv_cache()
        def run():
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            lg = model.llma.forward_inference(prompts, 0, img)
            e[1].record()
            ids = []
            nxt = lg.argmax(-1, keepdim=True)
            for i in range(G - 1):
                ids.append(nxt)
                lg = model.llma.forward_inference(nxt, 256 + i)
                nxt = lg.argmax(-1, keepdim=True)
            ids.append(nxt)
            e[2].record(); torch.cuda.synchronize()
            return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]) / max(G - 1, 1), torch.cat(ids, 1)
        run()
        pre, dec, ids = run()
        same_rows = bool((ids == ids[:1]).all())
        if B == 1 and G == 128:
            ref_ids = ids[0, :32].tolist()
        match = None if ref_ids is None else sum(int(a == b) for a, b in zip(ids[0, :32].tolist(), ref_ids))
        total = pre + dec * (G - 1)
        rows.append({"image": S, "batch": B, "decode": G, "image_words": eng.image_words, "prefill_ms": pre, "decode_ms_per_token": dec,
                     "images_per_s": B * 1e3 / total, "tokens_per_s": B * G * 1e3 / total, "rows_identical": same_rows,
                     "first32_ids_equal_to_batch1": match})
        print(rows[-1], flush=True)
    del model, eng
    torch.cuda.empty_cache()
json.dump({"config": "llama-7b + ViT-L/14 random init, 256-token prompt, greedy, 1xB200", "rows": rows}, open("gpurun_out/sweep%s.json" % ("_" + "_".join(sys.argv[1:]) if len(sys.argv) > 1 else ""), "w"), indent=1)
