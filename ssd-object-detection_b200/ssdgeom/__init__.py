"""ssdgeom -- B200-native (sm_100a) SSD box-geometry hot path behind the reference's own
function names.

    from ssdgeom.utils.bbox import iou, iou_n, match_bbox, apply_anchor_box      # utils/bbox.py
    from ssdgeom.models.ssd_model import SSDBoxGeometry, build_prior_box, ssd_loss  # models/ssd_model.py

Every call runs hand-written CUDA through libssdgeom.so (include/ssdgeom.h); there is no CPU
fallback and no other backend."""
from . import _native  # noqa: F401

__version__ = "0.1.0"
