"""BASELINE config 5: inference throughput sweep -- batch {1,4,16,64} x image {224,336,448} x decode {32,128,512},
Llama-2-7B + ViT-L/14 random init, 256-token prompts, greedy, through the public generation loop
(`MetaModel.generate_tokens`, the body of `generate`, reference M/eval_affordance_v2.py:331-339 / M/model/meta.py:379-485).

    python scripts/gpu_sweep_config5.py --gpus N [--cells quick]

Inference shards by replicas with no collective: N independent processes, one per GPU, run the
whole sweep CONCURRENTLY on the N GPUs of the box; per-cell images/s at N' <= N GPUs = sum over the first N'
replicas of that run (a conservative figure for N' < N: all N GPUs were loading the host / power envelope).
Replica 0 also times the CPU oracle (the reference's CPU path restated; 16 host cores on the GPU box) on one
cell -- batch 1, 336 px, 256-token prompt, 32 generated tokens -- with the SAME weights copied from the GPU."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, cells, cpu_cell):
    import torch

    import bench
    from a3vlm_b200.model import MetaModel

    cfg = bench.MODELS["7b"]
    dev = torch.device("cuda", 0)
    rows = []
    for S in sorted({c[1] for c in cells}):
        model = MetaModel("llama_b200", {**cfg, "image_size": S}, None, with_visual=True, max_seq_len=2048, max_batch_size=64)
        model.llma.reset_parameters(seed=2)
        g = torch.Generator().manual_seed(7)
        base_prompt = torch.randint(3, 32000, (256,), generator=g).tolist()
        base_prompt[0] = 1
        base_img = torch.randn(1, 3, S, S, generator=g)
        for B in sorted({c[0] for c in cells if c[1] == S}):
            prompts = [list(base_prompt) for _ in range(B)]
            img = base_img.repeat(B, 1, 1, 1).to(dev)
            model.llma.engine.destroy_kv_cache()
            model.generate_tokens(prompts, img, max_gen_len=8, eos_id=-1)  # warm-up: workspaces, decode graph of this batch
            for G in sorted({c[2] for c in cells if c[0] == B and c[1] == S}):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out, _, _ = model.generate_tokens(prompts, img, max_gen_len=G, eos_id=-1)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                assert all(len(o) == G for o in out)
                rows.append({"batch": B, "image": S, "decode": G, "image_words": model.llma.image_words, "seconds": dt,
                             "images_per_s": B / dt, "tokens_per_s": B * G / dt, "rows_identical": all(o == out[0] for o in out)})
                print(rank, rows[-1], flush=True)
        if rank == 0 and cpu_cell and S == 336:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import a3vlm_oracle as O

            torch.set_num_threads(os.cpu_count() or 1)
            sd = {k: t.detach().cpu() for k, t in model.llma.state_dict().items()}
            a = O.ModelArgs(**cfg, max_seq_len=2048, max_batch_size=1)
            orc = O.OracleTransformer(a, O.VisionArgs(image_size=336), sd, native_linear=True)
            t0 = time.perf_counter()
            ref, _, _ = O.generate_tokens(orc, [base_prompt], images=base_img, max_gen_len=32, eos_id=-1)
            dt = time.perf_counter() - t0
            got, _, _ = model.generate_tokens([base_prompt], base_img.to(dev), max_gen_len=32, eos_id=-1)
            same = next((i for i, (x, y) in enumerate(zip(got[0], ref[0])) if x != y), 32)
            rows.append({"cpu_oracle": True, "batch": 1, "image": 336, "decode": 32, "seconds": dt, "images_per_s": 1 / dt,
                         "tokens_per_s": 32 / dt, "cores": os.cpu_count(), "greedy_prefix_equal_to_cuda": same,
                         "note": "oracle/a3vlm_oracle.py (torch eager, bf16 linears) on the host cores, same weights; ids equal "
                                 "to the CUDA path up to the first bf16 near-tie of random-weight logits"})
            print(rank, rows[-1], flush=True)
            del orc, sd
        del model
        torch.cuda.empty_cache()
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"sweep5_rank{rank}.json"), "w"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cells", default="full", choices=["full", "quick"])
    ap.add_argument("--worker", type=int, default=-1)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.cells == "full":
        cells = [(b, s, g) for b in (1, 4, 16, 64) for s in (224, 336, 448) for g in (32, 128, 512)]
    else:
        cells = [(b, 336, g) for b in (1, 16) for g in (32, 128)]
    if args.worker >= 0:
        return worker(args.worker, cells, not args.no_cpu)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    procs = []
    t0 = time.perf_counter()
    for r in range(args.gpus):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(r))
        cmd = [sys.executable, os.path.abspath(__file__), "--worker", str(r), "--cells", args.cells] + (["--no-cpu"] if args.no_cpu else [])
        procs.append(subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate()[0] for p in procs]
    wall = time.perf_counter() - t0
    for r, (p, lg) in enumerate(zip(procs, logs)):
        if p.returncode != 0:
            print(f"replica {r} failed:\n{lg[-3000:]}")
            sys.exit(1)
    per = [json.load(open(os.path.join(ROOT, "gpurun_out", f"sweep5_rank{r}.json"))) for r in range(args.gpus)]
    table = []
    for c in cells:
        vals = [next(x["images_per_s"] for x in rows if not x.get("cpu_oracle") and (x["batch"], x["image"], x["decode"]) == c) for rows in per]
        row = {"batch": c[0], "image": c[1], "decode": c[2], "images_per_s_by_replica": [round(v, 3) for v in vals]}
        for n in (1, 2, 4, 8):
            if n <= args.gpus:
                row[f"images_per_s_{n}gpu"] = round(sum(vals[:n]), 3)
        row["tokens_per_s_all"] = round(sum(vals) * c[2], 1)
        table.append(row)
    cpu = [x for x in per[0] if x.get("cpu_oracle")]
    out = {"config": "llama-7b + ViT-L/14 random init, 256-token prompt, greedy through MetaModel.generate_tokens, host wall clock per call "
                     "(prefill + decode + result read-back)", "n_gpus": args.gpus, "replicas": "independent processes, one per GPU, run concurrently",
           "wall_s": wall, "cells": table, "cpu_oracle": cpu[0] if cpu else None}
    path = os.path.join(ROOT, "gpurun_out", f"sweep_config5_{args.gpus}gpu.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "cells"}))
    for row in table:
        print(row)


if __name__ == "__main__":
    main()
