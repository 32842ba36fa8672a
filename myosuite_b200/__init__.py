"""myosuite_b200 -- B200-native batched musculoskeletal simulator for MyoSuite's env.step hot path.

    from myosuite_b200 import make
    env = make("myoHandPoseRandom-v0", num_envs=4096)      # batched, tensors in / tensors out
    env = make("myoElbowPose1D6MRandom-v0")                # single env, the reference's gym call shapes
"""
__version__ = "0.1.0"


def make(env_id, num_envs=None, **kwargs):
    from .vec_env import make as _make
    return _make(env_id, num_envs=num_envs, **kwargs)


def registered_ids():
    from .vec_env import registered_ids as _r
    return _r()
