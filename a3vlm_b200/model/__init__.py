"""Host-side mirror of the reference's model package (model/accessory/model): `MetaModel`, the
`Tokenizer` wrapper and the `LLM.llama_b200` plugin module (`ModelArgs`, `Transformer`)."""
from .meta import MetaModel  # noqa: F401
from .tokenizer import Tokenizer  # noqa: F401
