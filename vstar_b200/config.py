"""Architecture description of the VSM (V* visual search model) hot path.

Field defaults = the released `craigwu/seal_vsm_7b` checkpoint: Vicuna-7B + CLIP ViT-L/14 + OWL-ViT-B/16 +
SAM prompt/mask decoder (/root/reference/VisualSearch/model/VSM.py:74-140; /root/reference/visual_search.py:28-52).
"""
from dataclasses import dataclass, asdict

IMAGE_TOKEN_INDEX = -200  # /root/reference/VisualSearch/utils/utils.py:8


@dataclass
class VSMConfig:
    # Llama (Vicuna-7B defaults)
    hidden: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    intermediate: int = 11008
    vocab: int = 32004
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    # CLIP ViT-L/14
    clip_hidden: int = 1024
    clip_layers: int = 24
    clip_heads: int = 16
    clip_inter: int = 4096
    clip_image: int = 224
    clip_patch: int = 14
    clip_select_layer: int = -2
    # OWL-ViT-B/16
    owl_hidden: int = 768
    owl_layers: int = 12
    owl_heads: int = 12
    owl_inter: int = 3072
    owl_image: int = 768
    owl_patch: int = 16
    owl_query_dim: int = 512     # class_head out dim == det query dim (config.out_dim)
    # SAM prompt/mask decoder (fixed by VSM.py:92-112)
    sam_dim: int = 256
    sam_depth: int = 2
    sam_heads: int = 8
    sam_mlp: int = 2048
    vit_eps: float = 1e-5
    loc_token_idx: int = 32001

    @property
    def head_dim(self):
        return self.hidden // self.n_heads

    @property
    def clip_tokens(self):
        return (self.clip_image // self.clip_patch) ** 2

    @property
    def owl_grid(self):
        return self.owl_image // self.owl_patch

    def to_dict(self):
        return asdict(self)


def tiny_config(**over) -> VSMConfig:
    """The reduced-width config used for the committed golden vectors: real
    head dims (128 Llama, 64 ViT), real token counts (257 / 2305 / 48x48), real
    SAM decoder; few layers and narrow hidden sizes."""
    kw = dict(hidden=256, n_layers=2, n_heads=2, intermediate=512, vocab=320,
              clip_hidden=128, clip_layers=3, clip_heads=2, clip_inter=256,
              owl_hidden=128, owl_layers=2, owl_heads=2, owl_inter=256,
              owl_query_dim=64, loc_token_idx=300)
    kw.update(over)
    return VSMConfig(**kw)


