"""pytorch_superpoint_b200 -- B200-native (sm_100a) SuperPoint homography-adaptation / extraction hot path.

Host-side mirror of the reference's interface for this path (``utils/utils.py``, ``export.py``,
``models/model_wrap.py``, ``models/SuperPointNet*.py`` of eric-yyjau/pytorch-superpoint) over the C ABI
of ``libspb200.so`` (``include/spb200.h``).  Importing the package needs neither a GPU nor the built
library; calling any operator does, and raises otherwise -- there is no CPU fallback.
"""
from .homographies import (EXPORT_PARAMS, make_ha_homographies, sample_ha_homographies_device,  # noqa: F401
                           sample_homography)
from .net import SuperPointNet_b200, SuperPointNet_pretrained_b200, fold_state_dict  # noqa: F401
from .ops import (combine_heatmap, compute_valid_mask, flattenDetection, getPtsFromHeatmap,  # noqa: F401
                  inv_warp_image_batch, nn_match_two_way, sample_desc_from_points)

from .tracker import PointTracker_b200  # noqa: F401

__version__ = "0.2.0"
