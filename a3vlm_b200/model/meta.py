"""`MetaModel` with the reference's surface (model/accessory/model/meta.py): construction from
`llama_type` + json configs + tokenizer, `forward(examples, labels, images) -> (loss, dict)`,
`generate`, `stream_generate`, `compute_logits`, `evaluate_examples`, `sample_top_p`,
`get_trainable_params`, `get_image_words`.  The host-side control flow (pad trimming, prompt
force-feeding, stop sequences, left truncation) is restated from the reference line by line; every
tensor op underneath runs in the sm_100a kernels of the `llama_b200` plugin.
"""
from __future__ import annotations

import importlib
import json
from typing import Dict, Iterable, List, Optional, Sequence, Union

import torch
import torch.nn as nn


def plan_generation(prompt_tokens: Sequence[Sequence[int]], max_seq_len: int, image_words: int, max_gen_len: int,
                    eos_id: int, stop_sequences: Iterable[Sequence[int]] = ()) -> Dict:
    """Host-side set-up of `generate` (meta.py:419-442), device-free: the sequence budget
    (`max_seq_len -= image_words` when images are given :423-425, `total_len` :427), left truncation of
    long prompts (:428-429), the pad-0 token matrix with the prompts written in (:431-435 --
    `input_text_mask[b, c]` is `c < prompt_len[b]`), `start_pos = min prompt length` (:436), and the stop
    sequences ([eos] first, :439-442) flattened for `a3_generate_step`."""
    prompts = [list(t) for t in prompt_tokens]
    min_prompt_size = min(len(t) for t in prompts)
    max_prompt_size = max(len(t) for t in prompts)
    max_seq_len = max_seq_len - image_words
    total_len = min(max_seq_len, max_gen_len + max_prompt_size)
    prompts = [t[-(max_seq_len - max_gen_len):] for t in prompts]
    tokens = torch.zeros((len(prompts), total_len), dtype=torch.long)
    for k, t in enumerate(prompts):
        tokens[k, :len(t)] = torch.tensor(t, dtype=torch.long)
    stops = [[eos_id]] + [list(s) for s in stop_sequences]
    offs = [0]
    for st in stops:
        offs.append(offs[-1] + len(st))
    return {"prompts": prompts, "total_len": total_len, "start_pos": min_prompt_size, "tokens": tokens,
            "prompt_len": torch.tensor([len(t) for t in prompts], dtype=torch.int32), "stops": stops,
            "stop_flat": torch.tensor([t for st in stops for t in st], dtype=torch.long),
            "stop_offsets": torch.tensor(offs, dtype=torch.int32)}


class MetaModel(nn.Module):
    def __init__(self, llama_type: str = "llama_b200", llama_config: Union[str, dict, Sequence] = (),
                 tokenizer_path: Optional[str] = None, with_visual: bool = False, max_seq_len: int = 4096,
                 pretrain_stage: bool = False, max_batch_size: int = 32) -> None:
        """meta.py:21-86.  `llama_config`: json path(s) and / or dicts merged left to right (:34-39).
        `tokenizer_path=None` is an extension for token-level use (synthetic benchmarks): then
        `vocab_size` must come from the config and the string entry points raise.  `max_batch_size`
        is 32 in the reference (:41); it is a parameter here because BASELINE's sweep goes to 64."""
        super().__init__()
        self.llama_type = llama_type
        self.with_visual = with_visual
        module = importlib.import_module(f"{__package__}.LLM.{llama_type}")  # the plugin seam, meta.py:30-32
        ModelArgs, Transformer = module.ModelArgs, module.Transformer
        llama_args: Dict = {}
        if isinstance(llama_config, (str, dict)):
            llama_config = [llama_config]
        for c in llama_config:
            if isinstance(c, dict):
                llama_args.update(c)
            else:
                with open(c, "r") as f:
                    llama_args.update(json.loads(f.read()))
        llama_args["max_seq_len"] = max_seq_len
        llama_args["max_batch_size"] = max_batch_size
        if tokenizer_path is not None:
            from .tokenizer import Tokenizer

            self.tokenizer = Tokenizer(model_path=tokenizer_path)
            llama_args["vocab_size"] = self.tokenizer.n_words
        else:
            self.tokenizer = None
            if llama_args.get("vocab_size", -1) <= 0:
                raise ValueError("without a tokenizer the config must carry vocab_size")
        args = ModelArgs(**llama_args)
        self.llma = Transformer(args, with_visual=with_visual)
        self.criterion = torch.nn.CrossEntropyLoss(ignore_index=0)
        self._set_default_trainability(pretrain_stage)
        self.is_peft = getattr(self.llma, "is_peft", False)

    @classmethod
    def from_pretrained(cls, pretrained_path: Union[str, List[str]], llama_type: Optional[str] = None,
                        llama_config: Optional[Union[str, List[str]]] = None, tokenizer_path: Optional[str] = None,
                        with_visual: bool = False, max_seq_len: int = 4096, mp_group=None,
                        dtype=torch.bfloat16, device="cuda", quant: bool = False, max_batch_size: int = 32):
        """meta.py:88-222: build the model from a checkpoint directory (or a list applied in order) --
        `llama_type` from its `meta.json`, the config from `config.json`, the tokenizer probed next to the
        weights -- then stream the `consolidated.*.pth` shards into the flat HBM buffers
        (a3vlm_b200.checkpoint: any model-parallel shard count is merged on the way in).
        Unsupported here, raising instead of silently differing: `hf://` paths (no network), a
        model-parallel group (this implementation is model-parallel size 1), 4-bit `quant`, a dtype other
        than bf16.  A `meta.json` naming one of the reference's own plugins is mapped to its B200
        counterpart: llama_ens5 -> llama_b200, llama_ens5_2images -> llama_b200_2images."""
        import os
        import warnings

        from .. import checkpoint as ckpt

        if isinstance(pretrained_path, str):
            pretrained_path = [pretrained_path]
        if pretrained_path is None or len(pretrained_path) == 0:
            raise ValueError("pretrained_path should be specified")
        for path in pretrained_path:
            if path.startswith("hf://"):
                raise NotImplementedError("hf:// checkpoints need network access; download them and pass the directory")
        if mp_group is not None:
            import torch.distributed as dist

            if dist.get_world_size(mp_group) != 1:
                raise NotImplementedError("llama_b200 is model-parallel size 1 (the weights fit one B200)")
        if quant:
            raise NotImplementedError("4-bit quantisation is outside the B200 hot path (SURVEY section 8)")
        if dtype != torch.bfloat16:
            raise NotImplementedError("llama_b200 computes in bf16 (fp32 masters live in the optimiser)")
        last = pretrained_path[-1]
        if llama_type is None:
            mj = os.path.join(last, "meta.json")
            if not os.path.exists(mj):
                raise ValueError("Cannot determine llama_type")
            with open(mj, "r") as f:
                llama_type = json.load(f)["llama_type"]
        from_reference_plugin = llama_type in ("llama_ens5", "llama_ens5_2images")
        llama_type = {"llama_ens5": "llama_b200", "llama_ens5_2images": "llama_b200_2images"}.get(llama_type, llama_type)
        if llama_config is None:
            cj = os.path.join(last, "config.json")
            llama_config = [cj] if os.path.exists(cj) else []
        if isinstance(llama_config, str):
            llama_config = [llama_config]
        cfgs: List = []
        for c in llama_config:  # config.json written by the reference carries only its own ModelArgs fields
            if isinstance(c, str):
                with open(c, "r") as f:
                    c = json.loads(f.read())
            cfgs.append(c)
        for c in cfgs:  # max_seq_len / max_batch_size of a saved config are set by the arguments instead
            c.pop("max_seq_len", None)
            c.pop("max_batch_size", None)
        if tokenizer_path is None:
            tokenizer_path = ckpt.probe_tokenizer_path_from_pretrained(last)
            # extension: a checkpoint saved from a tokenizer-less (token-level) model carries vocab_size in
            # its config.json and loads without one; otherwise as the reference (meta.py:189-193)
            if tokenizer_path is None and not any(c.get("vocab_size", -1) > 0 for c in cfgs):
                raise FileNotFoundError("No tokenizer available")
        if from_reference_plugin:
            # a checkpoint written by the reference's own plugin: loadable only if it was built ViT-only
            ckpt.check_vision_branch(pretrained_path, 1024)
        with torch.cuda.device(torch.device(device) if str(device) != "cuda" else torch.cuda.current_device()):
            model = cls(llama_type, cfgs, tokenizer_path, with_visual, max_seq_len, max_batch_size=max_batch_size)
        print(f"Loading pretrained weights from {pretrained_path} ...")
        load_result = ckpt.load_tensor_parallel_model_list(model, pretrained_path)
        if load_result != {"missing_keys": [], "unexpected_keys": []}:
            warnings.warn(f"checkpoint and model mismatch: \n{load_result}")
        else:
            print("all params match perfectly!")
        model.eval()
        return model

    # ------------------------------------------------------------------ trainability (meta.py:224-232)
    def get_trainable_params(self, pretrain_stage: bool = False):
        inner = self.llma.get_trainable_params(pretrain_stage)
        return {"llma." + n: p for n, p in inner.items()}

    def _set_default_trainability(self, pretrain_stage: bool = False):
        for _, v in self.named_parameters():
            v.requires_grad = False
        for _, v in self.get_trainable_params(pretrain_stage).items():
            v.requires_grad = True

    # ------------------------------------------------------------------ training forward (meta.py:234-263)
    def forward(self, examples, labels, images=None, depth_imgs=None):
        with torch.no_grad():
            non_zero_ = torch.count_nonzero(labels, dim=0)
            pos = non_zero_.shape[0] - 1
            nz = non_zero_.tolist()  # one host sync (the reference syncs once per probed column)
            while pos >= 0 and nz[pos] == 0:
                pos -= 1
            if pos == -1:  # nothing to predict in the whole batch
                print("nothing to predict in the whole batch!")
                pos = 2
            examples = examples[:, :pos + 1]
            labels = labels[:, :pos + 1]
        if hasattr(self.llma, "loss_forward"):
            # fused LM-head + shifted CE(ignore_index=0): same value as :251-262, logits stay on chip
            c_loss = self.llma.loss_forward(examples, labels, images)
            return c_loss, {}
        output = self.llma(examples, images)
        additional_loss = {}
        if isinstance(output, tuple):
            output, additional_loss = output
        output = output[:, :-1, :]
        labels = labels[:, 1:]
        if labels.sum() == 0:
            c_loss = output.mean() * 0
        else:
            c_loss = self.criterion(output.reshape(-1, output.shape[-1]), labels.flatten())
        return c_loss, additional_loss

    # ------------------------------------------------------------------ generation (meta.py:379-485)
    def _device(self):
        return next(self.parameters()).device

    @torch.inference_mode()
    def generate_tokens(self, prompt_tokens: Sequence[Sequence[int]], images=None, max_gen_len: int = 512,
                        temperature: float = 0.0, top_p: float = 0.95,
                        stop_sequences: Iterable[Sequence[int]] = (), eos_id: Optional[int] = None,
                        depth_images=None, generator: Optional[torch.Generator] = None, sync_lag: int = 4):
        """The token-level body of `generate` (meta.py:414-479).  Returns (generated ids per row,
        tokens tensor, stop_pos).

        Same control flow as the reference, with the per-token tensor work on the device kernels:
        the token choice is `a3_argmax` (temperature 0, :461) or `a3_sample_top_p` (:456-459,568-583,
        uniforms from `generator`), and the force-feed / token write / stop-sequence bookkeeping of
        :463-475 is ONE `a3_generate_step` launch that also publishes `stopped.all()` (:476) to a
        device flag.  The host reads that flag `sync_lag` steps late through pinned memory, so it never
        stalls the GPU on a per-token `.all()` sync; the (at most `sync_lag`) steps issued after every
        row had stopped change nothing the reference returns (stop_pos is frozen, their tokens are
        zeroed again below)."""
        from collections import deque

        from .. import ops

        dev = self._device()
        if images is not None:
            images = images.to(dev)
        if depth_images is not None:
            depth_images = depth_images.to(dev)
        bsz = len(prompt_tokens)
        args = self.llma.args
        assert bsz <= args.max_batch_size, (bsz, args.max_batch_size)
        if eos_id is None:
            eos_id = self.tokenizer.eos_id
        plan = plan_generation(prompt_tokens, args.max_seq_len, self.llma.image_words if images is not None else 0,
                               max_gen_len, eos_id, stop_sequences)
        prompt_tokens, total_len, l_stop_tokens = plan["prompts"], plan["total_len"], plan["stops"]
        tokens = plan["tokens"].to(dev)
        prompt_len = plan["prompt_len"].to(dev)
        stop_flat, stop_offs = plan["stop_flat"].to(dev), plan["stop_offsets"].to(dev)
        start_pos, prev_pos = plan["start_pos"], 0
        stopped = torch.zeros(bsz, dtype=torch.bool, device=dev)
        stop_pos = torch.full((bsz,), start_pos + 1, device=dev, dtype=torch.long)
        next_token = torch.empty(bsz, dtype=torch.long, device=dev)
        flags = torch.zeros(max(total_len, 1), dtype=torch.int32, device=dev)
        flags_h = torch.zeros(max(total_len, 1), dtype=torch.int32).pin_memory()
        uniforms = None
        if temperature > 0:
            uniforms = torch.rand((total_len, bsz), device=dev, dtype=torch.float32, generator=generator)
        pending = deque()
        last = None  # the step at which the reference's `if stopped.all(): break` fires
        for cur_pos in range(start_pos, total_len):
            first = prev_pos == 0
            if depth_images is not None:  # meta.py:447-451
                logits = self.llma.forward_inference(tokens[:, prev_pos:cur_pos], prev_pos,
                                                     images if first else None, depth_images if first else None)
            else:
                logits = self.llma.forward_inference(tokens[:, prev_pos:cur_pos], prev_pos, images if first else None)
            logits = logits.float().contiguous()  # already fp32 and contiguous: no kernel
            if temperature > 0:
                ops.sample_top_p(logits, temperature, top_p, uniforms[cur_pos], next_token)
            else:
                ops.argmax(logits, next_token)
            ops.generate_step(tokens, prompt_len, next_token, cur_pos, stop_flat, stop_offs, len(l_stop_tokens),
                              stopped, stop_pos, flags[cur_pos:cur_pos + 1])
            flags_h[cur_pos:cur_pos + 1].copy_(flags[cur_pos:cur_pos + 1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((cur_pos, ev))
            if len(pending) > sync_lag:
                p, e = pending.popleft()
                e.synchronize()
                if int(flags_h[p]):
                    last = p
                    break
            prev_pos = cur_pos
        while last is None and pending:
            p, e = pending.popleft()
            e.synchronize()
            if int(flags_h[p]):
                last = p
        if last is not None and last + 1 < total_len:
            tokens[:, last + 1:] = 0  # the reference stopped writing here
        rows = tokens.tolist()
        sp = stop_pos.tolist()
        out = [rows[i][len(prompt_tokens[i]):sp[i]] for i in range(bsz)]
        return out, tokens, stop_pos

    def _argmax(self, logits: torch.Tensor) -> torch.Tensor:
        """Greedy choice (meta.py:461) on the device kernel: lowest index wins ties like torch.argmax."""
        from .. import ops

        out = torch.empty(logits.shape[0], device=logits.device, dtype=torch.int64)
        return ops.argmax(logits.contiguous(), out)

    @torch.inference_mode()
    def generate(self, prompts: List[str], images=None, depth_images=None, max_gen_len: int = 512,
                 temperature: float = 0.0, top_p: float = 0.95,
                 additional_stop_symbols: Iterable[str] = ()) -> List[str]:
        if isinstance(prompts, str):
            raise ValueError(f"{self.__class__}.generate expects a batched LIST of prompts, but str is given")
        if self.tokenizer is None:
            raise RuntimeError("generate(prompts: List[str]) needs a tokenizer; use generate_tokens")
        prompt_tokens = [self.tokenizer.encode(x, bos=True, eos=False) for x in prompts]
        stops = [self.tokenizer.encode_segment(s) for s in additional_stop_symbols]
        stops += [self.tokenizer.encode_wo_prefix_space(s) for s in additional_stop_symbols]
        out, _, _ = self.generate_tokens(prompt_tokens, images, max_gen_len, temperature, top_p, stops,
                                         depth_images=depth_images)
        return [self.tokenizer.decode(t) for t in out]

    @torch.inference_mode()
    def stream_generate(self, prompt: str, image=None, max_gen_len: int = 512, temperature: float = 0.0,
                        top_p: float = 0.95, additional_stop_symbols: Iterable[str] = ()):
        """meta.py:487-566: batch-1 generator yielding {"text", "end_of_content"}."""
        dev = self._device()
        args = self.llma.args
        prompt_tokens = self.tokenizer.encode(prompt, bos=True, eos=False)
        max_seq_len = args.max_seq_len
        if image is not None:
            max_seq_len -= self.llma.image_words
            if image.dim() == 4:
                assert image.shape[0] == 1
            else:
                assert image.dim() == 3
                image = image.unsqueeze(0)
            image = image.to(dev)
        max_prompt_size = max_seq_len - max_gen_len
        prompt_tokens = prompt_tokens[-max_prompt_size:]
        prompt_size = len(prompt_tokens)
        total_len = min(max_seq_len, max_gen_len + prompt_size)
        tokens = torch.full([total_len], 0, device=dev, dtype=torch.long)
        tokens[:prompt_size] = torch.tensor(prompt_tokens, device=dev).long()
        start_pos, prev_pos, generate_until = prompt_size, 0, prompt_size
        generated = ""
        for cur_pos in range(start_pos, total_len):
            logits = self.llma.forward_inference(tokens[None, prev_pos:cur_pos], prev_pos,
                                                 image if prev_pos == 0 else None).float()
            if temperature > 0:
                next_token = self.sample_top_p(torch.softmax(logits / temperature, dim=-1), top_p)
            else:
                next_token = self._argmax(logits)
            next_token = next_token.item()
            if next_token == self.tokenizer.eos_id:
                break
            tokens[cur_pos] = next_token
            prev_pos = cur_pos
            generate_until = cur_pos + 1
            generated = self.tokenizer.decode(tokens[start_pos:generate_until].tolist())
            stop = False
            for sym in additional_stop_symbols:
                idx = generated.find(sym)
                if idx != -1:
                    generated, stop = generated[:idx], True
                    break
            if stop:
                break
            yield {"text": generated, "end_of_content": False}
        yield {"text": generated, "end_of_content": True}

    def sample_top_p(self, probs, p):
        """meta.py:568-583, same signature (probabilities in).  `generate` itself draws with the fused
        `a3_sample_top_p` kernel straight from the logits; this torch form is kept for callers of the
        reference API."""
        probs_sort, probs_idx = torch.sort(probs, dim=-1, descending=True)
        probs_sum = torch.cumsum(probs_sort, dim=-1)
        mask = probs_sum - probs_sort > p
        probs_sort[mask] = 0.0
        probs_sort.div_(probs_sort.sum(dim=-1, keepdim=True))
        next_token = torch.multinomial(probs_sort, num_samples=1)
        return torch.gather(probs_idx, -1, next_token)

    # ------------------------------------------------------------------ scoring API (meta.py:265-377)
    @torch.inference_mode()
    def compute_logits(self, examples: List[Union[str, List[int]]], images=None, bos=True, eos=False):
        """meta.py:265-304: fp32 (T_i, V) logits per example of a right-padded (pad id 0) batch."""
        if isinstance(examples, str):
            raise ValueError(f"{self.__class__}.compute_logits expects a batched LIST of examples, but str is given")
        if isinstance(examples[0], str):
            examples = [self.tokenizer.encode(x, bos, eos) for x in examples]
        dev = self._device()
        if images is not None:
            images = images.to(dev)
        l_seq_len = [len(x) for x in examples]
        token_tensor = torch.full((len(examples), max(l_seq_len)), 0, device=dev, dtype=torch.long)
        for i, item in enumerate(examples):
            token_tensor[i, :len(item)] = torch.tensor(item, dtype=torch.long)
        output = self.llma(token_tensor, images)
        logits = output[0] if isinstance(output, tuple) else output
        return [lg[:n].float() for lg, n in zip(logits, l_seq_len)]

    @torch.inference_mode()
    def evaluate_examples(self, examples, contexts=None, images=None, bos=True, eos=False) -> Dict[str, List]:
        """meta.py:306-377: per example, over the part after its context, the log-likelihood, the reference's
        "ppl" (which is the MEAN token loss, not its exponential -- kept as is), whether greedy decoding would
        reproduce it (`max_equal`), and the logits themselves (`non_context_logits`)."""
        if isinstance(examples, str):
            raise ValueError(f"{self.__class__}.evaluate_examples expects a batched LIST of examples, but str is given")
        if isinstance(examples[0], str):
            examples = [self.tokenizer.encode(x, bos, eos) for x in examples]
            if contexts is not None:
                contexts = [self.tokenizer.encode(x, bos, False) for x in contexts]
        if contexts is not None:  # every example must start with its context
            assert all(list(e[:len(c)]) == list(c) for e, c in zip(examples, contexts))
        logits = self.compute_logits(examples, images)
        loss_func = torch.nn.CrossEntropyLoss(reduction="none", ignore_index=0)
        result: Dict[str, List] = {"log_likelihood": [], "ppl": [], "max_equal": [], "non_context_logits": []}
        for idx, item_logits in enumerate(logits):
            start = 0 if contexts is None else len(contexts[idx]) - 1
            assert start >= 0
            item_logits = item_logits[start:-1]
            labels = torch.tensor(list(examples[idx][start + 1:]), dtype=torch.long, device=item_logits.device)
            loss = loss_func(item_logits, labels)
            result["log_likelihood"].append(-loss.sum().item())
            result["ppl"].append(loss.mean().item())
            result["max_equal"].append((item_logits.argmax(dim=-1) == labels).all().item())
            result["non_context_logits"].append(item_logits)
        return result

    def get_image_words(self):
        return self.llma.image_words

    def get_quant_blocklist(self) -> List[str]:
        if hasattr(self.llma, "get_quant_blocklist"):
            return ["llma." + x for x in self.llma.get_quant_blocklist()]
        return []

    def get_basic_block_classes(self):
        if hasattr(self.llma, "get_basic_block_classes"):
            return self.llma.get_basic_block_classes()
        return [type(self.llma.layers[0])]
