"""Stand-in for gym.envs.robotics.rotations (gym is not installed in this image; test infrastructure).  robogym carries the
same rotation helpers in robogym.utils.rotation (the rearrange code imports both), so this module re-exports the reference's
own functions at test time instead of restating them."""
from robogym.utils.rotation import (  # noqa: F401
    euler2mat,
    euler2quat,
    mat2euler,
    mat2quat,
    normalize_angles,
    quat2euler,
    quat2mat,
    quat_conjugate,
    quat_identity,
    quat_mul,
    quat_rot_vec,
    subtract_euler,
)
