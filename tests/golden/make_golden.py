"""Generates tests/golden/tiny_e2e.pt from the CPU oracle (oracle/a3vlm_oracle.py).

    python tests/golden/make_golden.py

The reference itself cannot be imported (absent in-tree modules, missing wheels),
so these vectors come from the oracle restatement, which is pinned against transformers' Llama and CLIP
implementations in tests/test_oracle.py.  The vectors travel to the GPU box (where neither
/root/reference nor a second opinion exists) and are what tests/test_e2e_gpu.py compares the CUDA
path with; tests/test_oracle.py::test_golden_vectors_reproduce guards the oracle against drift."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import a3vlm_oracle as O  # noqa: E402

SEED, OUTPUT_SCALE = 11, 8.0


def tiny_cfg():
    a = O.ModelArgs(dim=256, n_layers=2, n_heads=2, vocab_size=512, multiple_of=64, max_seq_len=64, max_batch_size=4)
    v = O.VisionArgs(image_size=28, patch=14, width=128, layers=2, heads=2, mlp=256)
    return a, v


def tiny_inputs():
    g = torch.Generator().manual_seed(5)
    B, T = 2, 9
    tokens = torch.randint(3, 512, (B, T), generator=g)
    tokens[:, 0] = 1
    labels = tokens.clone()
    labels[:, :4] = 0
    image = torch.randn(B, 3, 28, 28, generator=g)
    return tokens, labels, image


def main():
    a, v = tiny_cfg()
    tokens, labels, image = tiny_inputs()
    sd = O.init_params(a, v, seed=SEED, output_scale=OUTPUT_SCALE)
    model = O.OracleTransformer(a, v, sd)
    out = {"seed": SEED, "output_scale": OUTPUT_SCALE, "tokens": tokens, "labels": labels, "image": image}
    with torch.no_grad():
        out["image_tokens"] = model.encode_image(image)
        out["train_logits"] = model.forward(tokens, image)
        out["train_logits_noimg"] = model.forward(tokens, None)
        out["prefill_logits"] = model.forward_inference(tokens[:, :6], 0, image)
        out["decode_logits"] = model.forward_inference(tokens[:, 6:7], 6)
    # fine-tune step: loss, grads (autograd through the same rounding points), one clipped AdamW step
    train_names = [k for k in sd if not k.startswith("clip.")]
    for k in train_names:
        sd[k].requires_grad_(True)
    loss = O.meta_forward_loss(model, tokens, labels, image)
    loss.backward()
    out["loss"] = loss.detach()
    grads = {k: sd[k].grad.detach().clone() for k in train_names}
    out["grad_norm"] = {k: float(g.float().norm()) for k, g in grads.items()}
    out["grad_head"] = {k: g.flatten()[:64].clone() for k, g in grads.items()}
    full = ["layers.0.attention.wq.weight", "layers.1.ffn_norm.weight", "norm.weight", "visual_proj.0.bias",
            "visual_proj.1.weight", "start_img", "end_img"]
    out["grad_full"] = {k: grads[k] for k in full}
    for k in train_names:
        sd[k].requires_grad_(False)
    master = {k: sd[k].detach().float().clone() for k in train_names}
    state = {k: (torch.zeros_like(master[k]), torch.zeros_like(master[k])) for k in train_names}
    out["total_norm"] = O.adamw_step(master, grads, state, 1, 1e-3, 0.1, 0.5)
    out["post_step"] = {k: master[k].clone() for k in full}
    # greedy decode with unequal prompts (force-feeding) through the image path
    prompts = [[1, 17, 33, 49], [1, 100, 101, 102, 103, 104]]
    model2 = O.OracleTransformer(a, v, {k: t.detach() for k, t in sd.items()})
    gen, toks, stop_pos = O.generate_tokens(model2, prompts, images=image, max_gen_len=8, eos_id=2)
    out.update(prompts=prompts, max_gen_len=8, generated=gen, gen_tokens=toks)
    torch.save(out, os.path.join(HERE, "tiny_e2e.pt"))
    print("loss", float(loss), "total_norm", out["total_norm"], "generated", gen)
    lg = out["prefill_logits"]
    top2 = lg.topk(2, -1).values
    print("prefill top-2 margin", (top2[:, 0] - top2[:, 1]).tolist())


if __name__ == "__main__":
    main()
