"""a3vlm_b200 -- B200-native (sm_100a) implementation of A3VLM's multimodal forward / backward hot
path behind the reference's `llama_type` plugin API.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
