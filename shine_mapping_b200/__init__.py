"""shine_mapping_b200 — B200-native (sm_100a) implementation of SHINE-mapping's per-point SDF training step
behind the reference's own `FeatureOctree` / `Decoder` / `sdf_bce_loss` surfaces."""
from .config import SHINEConfig
from .decoder import Decoder
from .feature_octree import FeatureOctree
from .fused import sdf_bce_step, sdf_infer
from .loss import sdf_bce_loss
from .trainer import SdfTrainer

__all__ = ["SHINEConfig", "Decoder", "FeatureOctree", "sdf_bce_step", "sdf_infer", "sdf_bce_loss", "SdfTrainer"]
