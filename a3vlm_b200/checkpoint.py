"""Checkpoint load / save with the reference's on-disk formats and call shapes
(model/accessory/util/tensor_parallel.py, model/accessory/util/misc.py:324-560), built for this
engine's memory layout.

What is the same as the reference: the three formats (`meta_ori` = Meta's `consolidated.NN.pth`,
`consolidated` = `consolidated.NN-of-MM.model.pth` with an optional {"model": ...} wrapper and `llma.`
key prefix, `consolidated_diff` = deltas added onto what is already loaded), format / shard-count
inference from the file names, which weights are split along which dimension (fairscale
ColumnParallelLinear -> dim 0, RowParallelLinear -> dim 1, ParallelEmbedding -> dim 1;
tensor_parallel.py:35-39 with the layer classes of LLM/llama_ens5.py:63-90,202-217,258-269), the
unequal-replica warning, the `{"missing_keys", "unexpected_keys"}` result, the `epochN[-iterM]`
directory with `consolidated.00-of-01.{model,optimizer,other}.pth`, `config.json`, `meta.json`, the
tokenizer files and `rank-specific-*.pth`.

What is different (B200-first): nothing is concatenated on the host.  The model's parameters are views
of flat HBM buffers in which wq|wk|wv and w1|w3 of a layer are row ranges of one fused matrix
(engine.build_layout), so every tensor-parallel shard of every tensor is copied ONCE, from the
memory-mapped checkpoint file straight into its slice of the device-resident parameter -- the merge of
the model-parallel ranks and the re-layout into the kernel-side format are that one host-to-device copy.
Peak host memory is the page cache of the files; there is no merged CPU state dict and no second
`load_state_dict` pass.  The model-parallel world size of this implementation is always 1 (13.5 / 26 GB
of bf16 weights fit one 180 GB GPU), so only the reference's "merge ranks" direction exists.
"""
from __future__ import annotations

import dataclasses
import json
import os
import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

FORMAT_FILENAME_PATTERNS: Dict[str, "re.Pattern"] = {  # tensor_parallel.py:41-46
    "meta_ori": re.compile(r"^consolidated.(\d{2}).pth$"),
    "consolidated": re.compile(r"^consolidated.(\d{2})-of-(\d{2}).model.pth$"),
    "consolidated_diff": re.compile(r"^consolidated.(\d{2})-of-(\d{2}).model-diff.pth$"),
}

# Which dimension each weight of the path is split along in a model-parallel checkpoint.  The
# reference derives this from the module classes (tensor_parallel.py:35-39,49-59); the classes are
# fairscale's, absent here, so the rule is stated on the parameter names those classes carry in
# LLM/llama_ens5.py: wq / wk / wv / w1 / w3 / output are ColumnParallelLinear (:63-84,202-212,267-269),
# wo / w2 RowParallelLinear (:86-90,207-209), tok_embeddings ParallelEmbedding (:258-260).
_COLUMN = re.compile(r"(^|\.)(layers\.\d+\.attention\.w[qkv]|layers\.\d+\.feed_forward\.w[13]|output)\.weight$")
_ROW = re.compile(r"(^|\.)(layers\.\d+\.attention\.wo|layers\.\d+\.feed_forward\.w2)\.weight$")
_EMBED = re.compile(r"(^|\.)tok_embeddings\.weight$")


class ColumnParallelLinear(nn.Module):
    """Name-compatible stand-ins (model-parallel size 1) for the fairscale layer classes, so that the
    reference's own loader can be taught about this model with
    `tensor_parallel._MODEL_PARALLEL_MODULES += a3vlm_b200.checkpoint.MODEL_PARALLEL_MODULES`
    (the registration hook that file documents at :17-33)."""


class RowParallelLinear(nn.Module):
    pass


class ParallelEmbedding(nn.Module):
    pass


MODEL_PARALLEL_MODULES: List[Tuple[type, Dict[str, int]]] = [  # tensor_parallel.py:35-39
    (ColumnParallelLinear, {"weight": 0, "bias": 0}),
    (RowParallelLinear, {"weight": 1, "bias": -1}),
    (ParallelEmbedding, {"weight": 1}),
]


def holder_class_for(name: str):
    """The stand-in class the module owning parameter `name` should be an instance of (None: a
    replicated parameter)."""
    if _COLUMN.search(name):
        return ColumnParallelLinear
    if _ROW.search(name):
        return RowParallelLinear
    if _EMBED.search(name):
        return ParallelEmbedding
    return None


def weight_parallel_dim(name: str) -> Optional[int]:
    """Split dimension of parameter `name` (with or without the `llma.` prefix); None = replicated."""
    if _COLUMN.search(name):
        return 0
    if _ROW.search(name) or _EMBED.search(name):
        return 1
    return None


def get_weight_parallel_dim(model: Union[nn.Module, Iterable[str]]) -> Dict[str, int]:
    """tensor_parallel.py:49-59: {parameter name: split dim} for the tensor-parallel parameters."""
    names = [n for n, _ in model.named_parameters()] if isinstance(model, nn.Module) else list(model)
    return {n: d for n in names for d in [weight_parallel_dim(n)] if d is not None}


def get_tensor_parallel_shards_file_name(format: str, mp_size: int) -> List[str]:
    """tensor_parallel.py:172-199."""
    return {
        "meta_ori": [f"consolidated.{i:02d}.pth" for i in range(mp_size)],
        "consolidated": [f"consolidated.{i:02d}-of-{mp_size:02d}.model.pth" for i in range(mp_size)],
        "consolidated_diff": [f"consolidated.{i:02d}-of-{mp_size:02d}.model-diff.pth" for i in range(mp_size)],
    }[format]


def infer_checkpoint_format_and_mp_size(path: str) -> Tuple[str, int]:
    """tensor_parallel.py:333-386: format and model-parallel size from the file names; raises
    NotImplementedError on an unrecognised, mixed or incomplete folder like the reference."""
    if not os.path.isdir(path):
        raise NotImplementedError("The given path does not point to a valid folder.")
    files = [fn for fn in os.listdir(path) if os.path.isfile(os.path.join(path, fn))]
    fmt, mp = None, None
    for name, pattern in FORMAT_FILENAME_PATTERNS.items():
        matched = [fn for fn in files if pattern.match(fn)]
        if matched:
            if fmt is not None:
                raise NotImplementedError(f"Multiple matched format detected: {fmt} and {name}.")
            fmt, mp = name, len(matched)
    if fmt is None:
        raise NotImplementedError(f"Files in the given folder do not match any format. Contents in the folder: {sorted(os.listdir(path))}.")
    for fn in get_tensor_parallel_shards_file_name(fmt, mp):
        if fn not in files:
            raise NotImplementedError("An expected file is not found in the target folder: " + fn)
    return fmt, mp


def load_tensor_parallel_shard_state_dict(path: str, format: str, shard_id: int, num_shards: int,
                                          mmap: bool = True) -> Dict[str, torch.Tensor]:
    """tensor_parallel.py:202-229, memory-mapped when the file allows it (tensors are then paged in
    only as they are copied to the device)."""
    shard_path = os.path.join(path, get_tensor_parallel_shards_file_name(format, num_shards)[shard_id])
    try:
        shard = torch.load(shard_path, map_location="cpu", mmap=mmap, weights_only=False)
    except (RuntimeError, ValueError):  # legacy (non-zip) serialisation cannot be mapped
        shard = torch.load(shard_path, map_location="cpu", weights_only=False)
    if format.startswith("consolidated"):
        if "model" in shard and isinstance(shard["model"], dict):
            shard = shard["model"]
    elif format == "meta_ori":
        shard = {"llma." + k: v for k, v in shard.items()}
    return shard


def tensor_load_shard(target: torch.Tensor, parallel_dim: int, num_shards: int, shard_id: int,
                      value: torch.Tensor, mode: str = "set") -> None:
    """tensor_parallel.py:487-522: write shard `shard_id` of `num_shards` into its slice of `target`
    (parallel_dim -1 = replicated: only shard 0 is written).  `target` may live on the device: the value
    is converted and copied in one step."""
    assert parallel_dim < target.ndim or parallel_dim == -1
    if parallel_dim == -1:
        if shard_id != 0:
            return
        dst = target
    else:
        if target.size(parallel_dim) % num_shards:
            raise ValueError(f"dimension {parallel_dim} of size {target.size(parallel_dim)} does not split into {num_shards} shards")
        n = target.size(parallel_dim) // num_shards
        dst = target.narrow(parallel_dim, n * shard_id, n)
    if tuple(dst.shape) != tuple(value.shape):
        if dst.numel() != value.numel():
            raise ValueError(f"shape mismatch: parameter slice {tuple(dst.shape)} vs checkpoint {tuple(value.shape)}")
        value = value.reshape(dst.shape)
    if mode == "set":
        dst.copy_(value)  # dtype conversion + host-to-device in one copy
    elif mode == "add":
        dst.add_(value.to(device=dst.device, dtype=dst.dtype))
    else:
        raise NotImplementedError(f"Unknown mode: {mode}.")


@torch.no_grad()
def load_into(targets: Dict[str, torch.Tensor], path_list: Union[str, Sequence[str]], verbose: bool = False,
              strip_prefix: str = "") -> Dict[str, List[str]]:
    """Core of `load_tensor_parallel_model_list` (tensor_parallel.py:427-484) on a plain
    {name: tensor} mapping (device or host tensors -- the CPU tests drive this directly).  Checkpoints
    are applied in order; a base checkpoint overrides, a *_diff checkpoint adds to keys loaded before
    and sets the others (:388-424).  Shards of a split tensor go straight into their slice of the
    target; replicated tensors come from shard 0 and unequal replicas are reported like the reference
    (:118-126).  `strip_prefix`: remove this prefix (e.g. "llma.") from checkpoint keys first."""
    if isinstance(path_list, str):
        path_list = [path_list]
    existing: set = set()
    missing = set(targets.keys())
    unexpected: set = set()
    for i, path in enumerate(path_list):
        fmt, mp = infer_checkpoint_format_and_mp_size(path)
        print(f"Loading from checkpoint at: {path} ({i + 1} of {len(path_list)}, format is \"{fmt}\", {mp} shard(s))")
        assert i != 0 or not fmt.endswith("_diff"), "The first checkpoint in the list cannot be a *_diff checkpoint."
        shards = [load_tensor_parallel_shard_state_dict(path, fmt, r, mp) for r in range(mp)]
        seen_here = []
        for key in dict.fromkeys(k for sh in shards for k in sh):  # first-seen order, no duplicates
            name = key[len(strip_prefix):] if strip_prefix and key.startswith(strip_prefix) else key
            if name not in targets:
                print(f"discard unexpected parameter: {key}")  # tensor_parallel.py:108-110
                continue
            target = targets[name]
            pieces = [(r, sh[key]) for r, sh in enumerate(shards) if key in sh]
            mode = "add" if fmt.endswith("_diff") and name in existing else "set"
            dim = weight_parallel_dim(name)
            if dim is not None and len(pieces) == mp and mp > 1:
                for r, val in pieces:
                    tensor_load_shard(target, dim, mp, r, val, mode)
            else:
                if len(pieces) > 1:  # replicated parameter stored by every rank
                    ref = pieces[0][1]
                    for r, val in pieces[1:]:
                        if val.shape != ref.shape or not torch.equal(val, ref):
                            md = (val.float() - ref.float()).abs().max().item() if val.shape == ref.shape else float("nan")
                            print(f"WARNING! Found unequal replicas of non-tensor-parallel params: name={key}, "
                                  f"ranks={[q for q, _ in pieces]}, max_diff={md}.")
                            break
                tensor_load_shard(target, -1, 1, 0, pieces[0][1], mode)
            seen_here.append(name)
            if verbose:
                print(f"  {key}: {mode}, {len(pieces)} piece(s), dim {dim}")
        existing.update(seen_here)
        missing.difference_update(seen_here)
        del shards
    return {"missing_keys": sorted(missing), "unexpected_keys": sorted(unexpected)}


def check_vision_branch(path_list: Union[str, Sequence[str]], vit_width: int) -> None:
    """A released reference checkpoint (`llama_ens5` / `llama_ens5_2images`) carries the four-encoder ensemble:
    `qformer_proj.*`, ConvNeXt / DINOv2 towers and a 5632-wide `visual_proj.0.weight` (LLM/llama_ens5.py:303-333).
    The llama_b200 plugins build the CLIP ViT-L/14 branch only, so such a checkpoint cannot be
    mapped onto them: say so, instead of failing later with a generic shape mismatch."""
    if isinstance(path_list, str):
        path_list = [path_list]
    for path in path_list:
        fmt, mp = infer_checkpoint_format_and_mp_size(path)
        shard = load_tensor_parallel_shard_state_dict(path, fmt, 0, mp)
        bad = [k for k in shard if "qformer" in k or ".openclip_convnext_xxl." in k or ".dinov2_vitg14." in k]
        vp = next((v for k, v in shard.items() if k.endswith("visual_proj.0.weight")), None)
        if vp is not None and vp.dim() == 2 and vp.shape[1] != vit_width:
            bad.append(f"visual_proj.0.weight of width {vp.shape[1]} (this build: {vit_width})")
        if bad:
            raise NotImplementedError(
                f"{path}: checkpoint of the reference's four-encoder vision ensemble ({bad[0]}, {len(bad)} such entries); "
                "llama_b200 implements the CLIP ViT-L/14 + visual_proj + Llama path only -- the ConvNeXt-XXL / DINOv2 / "
                "Q-Former branches are out of scope (DESIGN.md section 7)")


def _inner(model):
    return getattr(model, "llma", model)


def load_tensor_parallel_model_list(model: nn.Module, path_list: Union[str, Sequence[str]],
                                    verbose: bool = False) -> Dict[str, List[str]]:
    """tensor_parallel.py:427-484 for a `MetaModel` (keys carry the `llma.` prefix) or a bare
    `Transformer`.  Streams every shard into the flat HBM buffers, then lets the engine re-derive its
    kernel-side copies (padded conv1, fp32 masters)."""
    inner = _inner(model)
    if hasattr(inner, "engine"):
        inner.engine.wait_weights()
    targets = dict(model.named_parameters())
    # checkpoint keys carry "llma." (MetaModel's attribute name); a bare Transformer's parameters do not
    result = load_into(targets, path_list, verbose, strip_prefix="llma." if inner is model else "")
    if hasattr(inner, "_check_alias"):
        inner._check_alias()
    if hasattr(inner, "engine"):
        inner.engine.weights_changed()
    return result


def load_tensor_parallel_model(model: nn.Module, path: str, format: str, verbose: bool = False):
    """tensor_parallel.py:299-330 (single path, explicit format)."""
    assert not format.endswith("_diff"), ("A *_diff checkpoint must be used together with the corresponding "
                                          "base checkpoint to obtain the full model weights.")
    fmt, _ = infer_checkpoint_format_and_mp_size(path)
    if fmt != format:
        raise NotImplementedError(f"\"{path}\" is not a valid {format} format checkpoint path")
    r = load_tensor_parallel_model_list(model, [path], verbose)
    print("load tensor parallel model result:\n", r)
    return r["missing_keys"], r["unexpected_keys"]


# ------------------------------------------------------------------------------------------ save / resume
_PRECISION = {"fp16": torch.float16, "bf16": torch.bfloat16, "tf32": torch.float}


def _rank_world():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def checkpoint_dir_name(epoch=None, iteration=None) -> str:
    name = f"epoch{epoch}"
    if iteration is not None:
        name += f"-iter{iteration}"
    return name


def save_checkpoint(output_dir: str, args, model, optimizer=None, loss_scaler=None, dataset_state=None,
                    epoch=None, iteration=None) -> str:
    """misc.py:324-438 with the same files: `consolidated.00-of-01.model.pth` ({"model": state dict with
    the reference's names, in args.precision}), tokenizer, `config.json` (= dataclasses.asdict of the
    inner ModelArgs), `meta.json`, `consolidated.00-of-01.optimizer.pth`, `...other.pth`, and one
    `rank-specific-RRRRR-of-WWWWW.pth` per rank.  Data-parallel rank 0 writes the shared files.  With the
    sharded optimiser every rank also writes its slice of the fp32 state
    (`optimizer-shard-RRRRR-of-WWWWW.pth`), since no rank holds all of it.  Returns the directory."""
    save_dir = os.path.join(output_dir, checkpoint_dir_name(epoch, iteration))
    os.makedirs(save_dir, exist_ok=True)
    rank, world = _rank_world()
    inner = _inner(model)
    dtype = _PRECISION[getattr(args, "precision", "bf16")] if args is not None else torch.bfloat16
    if rank == 0:
        sd = model.state_dict()
        if getattr(args, "only_save_trainable", False):
            keep = {".".join(p for p in k.split(".") if not p.startswith("_")) for k in model.get_trainable_params()}
            sd = {k: v for k, v in sd.items() if k in keep}
        torch.save({"model": {k: v.detach().to("cpu", dtype) for k, v in sd.items()}},
                   os.path.join(save_dir, "consolidated.00-of-01.model.pth"))
        tok = getattr(model, "tokenizer", None)
        if tok is not None:
            tok.save(save_dir)
        with open(os.path.join(save_dir, "config.json"), "w") as f:
            json.dump(dataclasses.asdict(inner.args), f, indent=2)
        with open(os.path.join(save_dir, "meta.json"), "w") as f:
            json.dump({"llama_type": getattr(model, "llama_type", "llama_b200")}, f, indent=2)
        torch.save({"epoch": epoch, "iter": iteration,
                    "scaler": loss_scaler.state_dict() if hasattr(loss_scaler, "state_dict") else loss_scaler,
                    "args": args}, os.path.join(save_dir, "consolidated.00-of-01.other.pth"))
    if optimizer is not None:
        osd = inner.engine.optimizer_state_dict()
        if osd["shard_world"] > 1:
            torch.save({"optimizer": osd}, os.path.join(save_dir, f"optimizer-shard-{rank:05d}-of-{world:05d}.pth"))
        elif rank == 0:
            torch.save({"optimizer": osd}, os.path.join(save_dir, "consolidated.00-of-01.optimizer.pth"))
    torch.save({"dataset_state": dataset_state},
               os.path.join(save_dir, f"rank-specific-{rank:05d}-of-{world:05d}.pth"))
    return save_dir


def get_latest_checkpoint_path(ckpt_path: str) -> Optional[str]:
    """misc.py:440-465: newest `epochN[-iterM]` sub-folder (a finished epoch outranks its iterations)."""
    it, ep = re.compile(r"^epoch(\d+)-iter(\d+)$"), re.compile(r"^epoch(\d+)$")
    parsed = []
    for folder in os.listdir(ckpt_path):
        if not os.path.isdir(os.path.join(ckpt_path, folder)):
            continue
        m, e = it.match(folder), ep.match(folder)
        if m:
            parsed.append((int(m.group(1)), int(m.group(2)), folder))
        elif e:
            parsed.append((int(e.group(1)), float("inf"), folder))
    parsed.sort(key=lambda x: (x[0], x[1]), reverse=True)
    return parsed[0][2] if parsed else None


def resume(resume_path: str, model, optimizer=None, loss_scaler=None, dataset_train=None):
    """misc.py:467-560 (resume_stage1 + resume_stage2 in one call -- there is no FSDP wrapping between
    them here): model weights, optimiser state, scaler, epoch / iteration, per-rank dataset state.
    Returns [epoch, iteration] to continue from, as the reference's `_load_other`."""
    if "epoch" not in os.path.basename(os.path.normpath(resume_path)):
        latest = get_latest_checkpoint_path(resume_path)
        if latest is None:
            print(f"no valid checkpoint found in {resume_path}")
            return None
        resume_path = os.path.join(resume_path, latest)
    print("Resume checkpoint %s" % resume_path)
    rank, world = _rank_world()
    inner = _inner(model)
    print("load model:", load_tensor_parallel_model_list(model, [resume_path]))
    if optimizer is not None:
        shard = os.path.join(resume_path, f"optimizer-shard-{rank:05d}-of-{world:05d}.pth")
        full = os.path.join(resume_path, "consolidated.00-of-01.optimizer.pth")
        osd = torch.load(shard if os.path.exists(shard) else full, map_location="cpu", weights_only=False)["optimizer"]
        inner.engine.load_optimizer_state_dict(osd)
    other = torch.load(os.path.join(resume_path, "consolidated.00-of-01.other.pth"), map_location="cpu",
                       weights_only=False)
    if loss_scaler is not None and other.get("scaler") is not None and hasattr(loss_scaler, "load_state_dict"):
        loss_scaler.load_state_dict(other["scaler"])
    epoch_iter = [int(other["epoch"]) if other.get("epoch") is not None else None,
                  int(other["iter"]) + 1 if other.get("iter") is not None else None]
    if epoch_iter[1] is None and epoch_iter[0] is not None:  # the epoch has finished
        epoch_iter[0] += 1
    rs = os.path.join(resume_path, f"rank-specific-{rank:05d}-of-{world:05d}.pth")
    if dataset_train is not None and os.path.exists(rs):
        try:
            dataset_train.load_state_dict(torch.load(rs, map_location="cpu", weights_only=False)["dataset_state"])
        except Exception as e:  # misc.py:551-556: a changed world size only loses the sampler position
            print(f"dataset state not restored: {e}")
    return epoch_iter


def probe_tokenizer_path_from_pretrained(pretrained_path: str) -> Optional[str]:
    """model/tokenizer.py:133-151: spm `tokenizer.model` first, then a HF `tokenizer.json` +
    `tokenizer_config.json` pair."""
    p = os.path.join(pretrained_path, "tokenizer.model")
    if os.path.exists(p):
        return p
    if os.path.exists(os.path.join(pretrained_path, "tokenizer.json")) and \
            os.path.exists(os.path.join(pretrained_path, "tokenizer_config.json")):
        return pretrained_path
    return None
