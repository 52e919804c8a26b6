"""Stand-in for gym.envs.robotics.utils (gym is not installed in this image): robogym's rearrange envs import it for
the mocap helpers.  The mocap path is outside this round's scope, so the helpers refuse to run."""


def _unsupported(*_a, **_k):
    raise NotImplementedError("gym.envs.robotics.utils mocap helpers are not available (mocap bodies are not supported)")


reset_mocap_welds = reset_mocap2body_xpos = mocap_set_action = ctrl_set_action = robot_get_obs = _unsupported
