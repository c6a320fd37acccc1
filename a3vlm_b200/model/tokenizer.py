"""Tokenizer front end with the reference's interface (model/accessory/model/tokenizer.py:15-131) for
standalone use of `a3vlm_b200.model.MetaModel`: a SentencePiece `.model` file or a Hugging Face tokenizer
directory behind `encode(s, bos, eos)`, `decode`, `encode_segment`, `encode_wo_prefix_space`, `n_words`,
`bos_id`, `eos_id`, `save`.  CPU string work -- not on the accelerated path (BASELINE.json: "the (B,A,S)
tokenizer/postprocess ... stay as-is"); inside the reference tree its own class is used unchanged."""
from __future__ import annotations

import os
from typing import Callable, List

# characters that tokenise on their own: text glued behind one of them starts without a word-initial space
_GLUE_PROBES = ("@", "\n", "\\", "=", ">", "`")


class _SentencePieceBackend:
    kind = "spm"

    def __init__(self, path: str):
        from sentencepiece import SentencePieceProcessor

        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        self.sp = SentencePieceProcessor(model_file=path)
        if self.sp.vocab_size() != self.sp.get_piece_size():
            raise ValueError("sentencepiece model: vocab_size != piece_size")
        self.bos_id, self.eos_id = self.sp.bos_id(), self.sp.eos_id()

    def ids(self, text: str) -> List[int]:
        return self.sp.encode(text)

    def text(self, ids: List[int]) -> str:
        return self.sp.decode(ids)

    def size(self) -> int:
        return self.sp.vocab_size()

    def save(self, directory: str) -> None:
        with open(os.path.join(directory, "tokenizer.model"), "wb") as f:
            f.write(self.sp.serialized_model_proto())


class _HuggingFaceBackend:
    kind = "transformers"

    def __init__(self, path: str):
        from transformers import AutoTokenizer

        self.hf = AutoTokenizer.from_pretrained(path, trust_remote_code=True)
        self.eos_id = self.hf.eos_token_id
        if self.eos_id is None:
            raise ValueError("the tokenizer defines no eos token")
        self.bos_id = self.hf.bos_token_id if self.hf.bos_token_id is not None else self.eos_id

    def ids(self, text: str) -> List[int]:
        return self.hf.encode(text, truncation=False, add_special_tokens=False)

    def text(self, ids: List[int]) -> str:
        return self.hf.decode(ids)

    def size(self) -> int:
        return len(self.hf)

    def save(self, directory: str) -> None:
        self.hf.save_pretrained(directory)


def _segment_needs_explicit_space(ids: Callable[[str], List[int]]) -> bool:
    """Does a mid-sentence segment have to be written with its leading space (tokenizer.py:92-114)?  Llama's
    SentencePiece model implies a word-initial space (False); byte-level BPE vocabularies do not (True)."""
    whole, bare = ids("Hi my darling"), ids("my darling")
    if whole[-len(bare):] == bare:
        return False
    spaced = ids(" my darling")
    if whole[-len(spaced):] != spaced:
        raise ValueError("cannot tell how this tokenizer treats word-initial spaces")
    return True


class Tokenizer:
    def __init__(self, model_path: str):
        backend = _SentencePieceBackend if model_path.endswith(".model") else _HuggingFaceBackend
        self._b = backend(model_path)
        self.tokenizer_type = self._b.kind
        self.tokenizer = getattr(self._b, "sp", None) or getattr(self._b, "hf")
        self.bos_id: int = self._b.bos_id
        self.eos_id: int = self._b.eos_id
        self.need_space_before_segment = _segment_needs_explicit_space(self._b.ids)

    def encode(self, s: str, bos: bool, eos: bool) -> List[int]:
        if type(s) is not str:
            raise TypeError("Tokenizer.encode expects a str")
        ids = self._b.ids(s)
        return [self.bos_id] * bool(bos) + ids + [self.eos_id] * bool(eos)

    def encode_segment(self, s: str) -> List[int]:
        """Token ids of `s` as it appears in the middle of a text, after a space."""
        s = s.lstrip(" ")
        return self._b.ids(" " + s if self.need_space_before_segment else s)

    def encode_wo_prefix_space(self, s: str) -> List[int]:
        """Token ids of `s` as it appears glued to the preceding character (no space in between)."""
        if self.need_space_before_segment:
            return self._b.ids(s)
        for probe in _GLUE_PROBES:
            head, both = self._b.ids(probe), self._b.ids(probe + s)
            if both[:len(head)] == head:
                return both[len(head):]
        raise NotImplementedError(f"every probe prefix merges into {s!r} during tokenisation")

    def decode(self, t: List[int]) -> str:
        return self._b.text(t)

    def save(self, save_dir: str) -> None:
        self._b.save(save_dir)

    @property
    def n_words(self) -> int:
        return self._b.size()
