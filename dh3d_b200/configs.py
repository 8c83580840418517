"""The handful of reference config fields that shape the forward pass (core/configs.py:35-144,
models/{local,global}/config.json); everything training-related is out of scope."""
from dataclasses import dataclass, field
from typing import List


@dataclass(frozen=True)
class DH3DConfig:
    num_points: int = 8192
    knn_num: int = 8
    init_feat_dim: int = 32
    featdim: int = 128
    dilate: int = 8              # backbone_local_dilate's dilate2 (always 8 in the reference)
    detection: bool = True       # detection_config / models/local/config.json
    extract_global: bool = True  # global_config / models/global/config.json
    gl_dilate: int = 8
    gl_dims: List[int] = field(default_factory=lambda: [256])
    cluster_size: int = 64       # global_netvald_block defaults
    output_dim: int = 256
    add_se: str = "max_pool"     # flex_conv_dilate's squeeze/excite pooling: 'max_pool' | 'avg_pool' | ''
    global_subsample: int = -1   # > 0: FPS-subsample the global-branch features before attention / NetVLAD

    @property
    def input_knn_indices(self):
        """The reference feeds CPU k-NN indices above 8192 points (core/model.py:38).  Our kernel
        has no such cap, so this only documents the reference's switch."""
        return self.num_points > 8192


def basic_config(**kw):
    return DH3DConfig(detection=False, extract_global=False, **kw)


def detection_config(**kw):
    return DH3DConfig(detection=True, extract_global=False, **kw)


def global_config(**kw):
    return DH3DConfig(detection=False, extract_global=True, **kw)


def full_config(**kw):
    """BASELINE.json config 3: local + detector + global in one pass (one shared backbone)."""
    return DH3DConfig(detection=True, extract_global=True, **kw)
