"""Drop-in proof, first half (CPU, this container): the reference's UNMODIFIED `MetaModel`
(/root/reference/model/accessory/model/meta.py, executed from where it lies; nothing of it is copied)
is constructed with `llama_type="llama_b200"` through its own `importlib` seam (meta.py:30-32) and its
own `__init__` / `forward` / `generate` / `get_image_words` / `get_trainable_params` drive the plugin.

    python tests/golden/make_seam_trace.py        # needs /root/reference -> tests/golden/reference_seam_trace.pt

There is no GPU in this container and the reference does not exist on the GPU box, so the proof is cut
at the seam itself:

  here     `accessory.model.LLM.llama_b200` is registered with OUR `ModelArgs` (the real dataclass) and
           a RECORDING DOUBLE of `Transformer` that honours the SURVEY 8b contract, answers every call
           with the CPU oracle and writes down exactly what the reference asked for: the constructor
           arguments the reference built from its JSON config, every `forward(examples, image)` and
           `forward_inference(tokens, start_pos, image)` call with its arguments, and what the
           reference made of the answers (loss, generated strings, image_words, trainable names).
  GPU box  tests/test_seam_gpu.py replays the recorded constructor arguments and call sequence against
           the REAL CUDA `llama_b200.Transformer` and checks every answer against the recorded one,
           then re-derives the reference's results (loss, strings) from the CUDA answers.

Together: same module name, same constructor call, same method calls, same results.  A third test
(tests/test_reference_pins.py::test_reference_metamodel_reaches_the_real_plugin, CPU) registers the
literal one-line stub of INTEGRATION.md and shows the reference's `__init__` reaching the real
constructor, which refuses to build without a CUDA device (no CPU fallback)."""
import dataclasses
import importlib
import json
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import a3vlm_oracle as O  # noqa: E402
import make_reference_golden as RG  # noqa: E402

SEED, OUTPUT_SCALE = 23, 8.0


def seam_cfg():
    """What the reference reads from its --llama_config JSON (meta.py:35-41 adds max_seq_len / max_batch_size /
    vocab_size itself)."""
    return dict(dim=256, n_layers=2, n_heads=2, multiple_of=64, norm_eps=1e-5, image_size=28, vit_width=128,
                vit_layers=2, vit_heads=2, vit_mlp=256)


def oracle_args(margs):
    a = O.ModelArgs(dim=margs.dim, n_layers=margs.n_layers, n_heads=margs.n_heads, vocab_size=margs.vocab_size,
                    multiple_of=margs.multiple_of, norm_eps=margs.norm_eps, max_seq_len=margs.max_seq_len,
                    max_batch_size=margs.max_batch_size)
    v = O.VisionArgs(image_size=margs.image_size, patch=margs.vit_patch, width=margs.vit_width, layers=margs.vit_layers,
                     heads=margs.vit_heads, mlp=margs.vit_mlp)
    return a, v


def install_seam(transformer_cls):
    """Stubs of make_reference_golden plus what MetaModel.__init__ itself touches; registers the plugin
    module under the name the reference imports."""
    RG.install_stubs()
    from a3vlm_b200.model.LLM import llama_b200 as ours

    class _Tok(RG.ToyTokenizer):
        def __init__(self, model_path=None):
            self.model_path = model_path

    sys.modules["accessory.model.tokenizer"].Tokenizer = _Tok
    sys.modules["accessory.util.misc"].mark_mp_params = lambda model: None
    mod = types.ModuleType("accessory.model.LLM.llama_b200")
    mod.ModelArgs, mod.Transformer = ours.ModelArgs, transformer_cls
    sys.modules["accessory.model.LLM.llama_b200"] = mod
    return importlib.import_module("accessory.model.meta")


def main():
    assert os.path.isdir(RG.REF), "needs the reference tree at /root/reference"
    trace = {"seed": SEED, "output_scale": OUTPUT_SCALE, "calls": []}

    class RecordingTransformer(nn.Module):
        """Contract of SURVEY 8b, answered by the CPU oracle; every call of the reference is written down."""

        def __init__(self, args, with_visual=False):
            super().__init__()
            trace["ctor"] = {"model_args": dataclasses.asdict(args), "with_visual": bool(with_visual),
                             "args_type": type(args).__module__ + "." + type(args).__name__}
            self.args = args
            a, v = oracle_args(args)
            sd = O.init_params(a, v if with_visual else None, seed=SEED, output_scale=OUTPUT_SCALE)
            self._orc = O.OracleTransformer(a, v if with_visual else None, sd)
            self.image_words = self._orc.image_words
            self.image_size = args.image_size
            self.cache_image_words = 0
            for k, t in sd.items():  # the reference walks named_parameters() and flips requires_grad
                holder = self
                parts = k.split(".")
                for p in parts[:-1]:
                    if p not in holder._modules:
                        holder.add_module(p, nn.Module())
                    holder = holder._modules[p]
                holder.register_parameter(parts[-1], nn.Parameter(t.clone(), requires_grad=False))
            self.layers = nn.ModuleList([self._modules["layers"]._modules[str(i)] for i in range(args.n_layers)])

        def get_trainable_params(self, pretrain_stage=False):
            return {n: p for n, p in self.named_parameters() if not n.startswith("clip.")}

        def forward(self, examples, image=None):
            with torch.no_grad():
                out = self._orc.forward(examples, image)
            trace["calls"].append({"fn": "forward", "examples": examples.clone(), "image": None if image is None else image.clone(),
                                   "logits": out.float().clone()})
            return out.float()

        def forward_inference(self, tokens, start_pos, image=None):
            out = self._orc.forward_inference(tokens, start_pos, image)
            self.cache_image_words = self._orc.cache_image_words
            top2 = out.float().topk(2, dim=-1).values
            trace["calls"].append({"fn": "forward_inference", "tokens": tokens.clone(), "start_pos": int(start_pos),
                                   "image": None if image is None else image.clone(), "logits": out.float().clone(),
                                   "margin": (top2[:, 0] - top2[:, 1]).clone()})
            return out

    M = install_seam(RecordingTransformer)
    with tempfile.TemporaryDirectory() as td:
        cfg_path = os.path.join(td, "params.json")
        json.dump(seam_cfg(), open(cfg_path, "w"))
        # --- the reference's own constructor, unmodified (meta.py:21-86)
        meta = M.MetaModel("llama_b200", cfg_path, os.path.join(td, "tokenizer.model"), with_visual=True, max_seq_len=64)
    trace["config_json"] = seam_cfg()
    trace["image_words"] = int(meta.get_image_words())
    trace["trainable"] = sorted(meta.get_trainable_params().keys())
    trace["requires_grad"] = {n: bool(p.requires_grad) for n, p in meta.named_parameters()}
    trace["basic_block"] = [c.__name__ for c in meta.get_basic_block_classes()]

    g = torch.Generator().manual_seed(5)
    ex = torch.randint(3, 512, (3, 12), generator=g)
    ex[:, 0] = 1
    lab = ex.clone()
    lab[:, :4] = 0
    ex[:, 9:], lab[:, 9:] = 0, 0  # padded tail: trimmed by the reference before the plugin sees it
    img = torch.randn(3, 3, 28, 28, generator=g)
    n0 = len(trace["calls"])
    with torch.no_grad():
        loss, extra = meta(ex, lab, images=img)  # meta.py:234-263
    assert extra == {} and len(trace["calls"]) == n0 + 1
    trace["forward_case"] = {"examples": ex, "labels": lab, "images": img, "loss": loss.clone(), "call": n0}

    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a_, **k_: self  # no GPU here; the reference's loop is device-agnostic
    try:
        n0 = len(trace["calls"])
        prompts = ["17 33 49", "100 101 102 103 104 105"]
        res = meta.generate(prompts, img[:2], None, 10, 0.0, 0.95, [])  # meta.py:379-485, greedy, with images
        trace["generate_case"] = {"prompts": prompts, "images": img[:2].clone(), "max_gen_len": 10, "result": list(res),
                                  "calls": (n0, len(trace["calls"]))}
        n0 = len(trace["calls"])
        res = meta.generate(["5 6", "9 10 11 12", "7"], None, None, 6, 0.0, 0.95, [])  # text only
        trace["generate_text_case"] = {"prompts": ["5 6", "9 10 11 12", "7"], "max_gen_len": 6, "result": list(res),
                                       "calls": (n0, len(trace["calls"]))}
    finally:
        torch.Tensor.cuda = cuda
    path = os.path.join(HERE, "reference_seam_trace.pt")
    torch.save(trace, path)
    print("wrote", path, "ctor", trace["ctor"]["model_args"], "calls", [(c["fn"], c.get("start_pos")) for c in trace["calls"]])
    print("loss", float(trace["forward_case"]["loss"]), "generated", trace["generate_case"]["result"], trace["generate_text_case"]["result"])
    print("min margin", min(float(c["margin"].min()) for c in trace["calls"] if c["fn"] == "forward_inference"))


if __name__ == "__main__":
    main()
