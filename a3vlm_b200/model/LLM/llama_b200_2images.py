"""`llama_type = "llama_b200_2images"`: the RGB-D build of the plugin, mirroring the reference's
LLM/llama_ens5_2images.py next to LLM/llama_ens5.py.  Differences from `llama_b200` (the diff of the two
reference files): `image_words` counts two image streams and `visual_image_words` one (:335-336), two more
parameters `start_depth_img` / `end_depth_img` (:343-344), and `forward_inference(tokens, start_pos, image,
depth_images)` splices BOS | RGB views | depth views | text when BOTH images are given (:511-547) -- the
depth image goes through the same CLIP tower and `visual_proj`.  Training is unchanged: the reference's
`MetaModel.forward` never passes the depth images on (M/model/meta.py:251)."""
from .llama_b200 import ModelArgs, Transformer as _Base, TransformerBlock  # noqa: F401


class Transformer(_Base):
    _RGBD = True
