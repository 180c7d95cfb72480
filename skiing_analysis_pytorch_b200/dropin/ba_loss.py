"""Stand-in for ``bundle_adjustment/loss.py``: the same eight public names, CUDA arithmetic
(skiing_analysis_pytorch_b200/losses.py -> libska.so), differentiable like the reference."""
from ..losses import (  # noqa: F401
    BONES,
    baseline_reg_loss,
    bone_length_loss,
    camera_center_from_Rt,
    camera_smooth_loss,
    pose_temporal_loss,
    project_points,
    reprojection_loss,
)
from ..ba import run_local_ba  # noqa: F401  the optimiser vggt/multi_view_process.py:553 expects
