"""The C-ABI boundary: libb2h.so builds for sm_100a, loads, and exports every symbol include/b2h.h declares.
No compute calls here (no GPU in this suite): b2h_create must refuse loudly instead of falling back to the CPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from mujocoposelearning_b200.build import build
    build()
    from mujocoposelearning_b200.lib import load
    return load()


def test_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "b2h.h").read_text()
    declared = set(re.findall(r"\b(b2h_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 18
    raw = C.CDLL(str(ROOT / "mujocoposelearning_b200" / "libb2h.so"))
    missing = [n for n in sorted(declared) if not hasattr(raw, n)]
    assert not missing, missing


def test_struct_layouts_match(lib):
    from mujocoposelearning_b200 import abi
    assert lib.b2h_abi_version() == 2      # 2: B2HConfig::sensor_terms, the rollout / packed-policy entry points
    assert lib.b2h_sizeof_model() == C.sizeof(abi.B2HModel) and lib.b2h_sizeof_config() == C.sizeof(abi.B2HConfig)


def test_create_fails_loudly_without_gpu(lib, model_struct):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mujocoposelearning_b200 import abi
    cfg = abi.make_config(4, frame_skip=3, duration=10.0)
    h = C.c_void_p()
    rc = lib.b2h_create(C.byref(model_struct), C.byref(cfg), C.byref(h))
    assert rc == abi.ECUDA and b"no CPU path" in lib.b2h_last_error()
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.lib import B2HError
    with pytest.raises(B2HError):
        HumanoidBatch(4)


def test_bad_arguments(lib, model_struct):
    from mujocoposelearning_b200 import abi
    cfg = abi.make_config(4)
    cfg.reward_type = 9
    h = C.c_void_p()
    assert lib.b2h_create(C.byref(model_struct), C.byref(cfg), C.byref(h)) == abi.EINVAL
    assert lib.b2h_gae(None, None, None, None, None, 0.99, 0.95, 4, 4, None, None, None) == abi.EINVAL


def test_launch_shape_rule(lib):
    """b2h_choose_launch_shape is pure host logic: B200 numbers (148 SMs, 227 KB opt-in shared memory per CTA)."""
    import ctypes as C

    def shape(n_envs, dtype=0):
        w, r = C.c_int(), C.c_int()
        assert lib.b2h_choose_launch_shape(n_envs, 148, 232448, dtype, C.byref(w), C.byref(r)) == 0
        return w.value, r.value
    # shared rows: as many as fit beside the model tables (12 KB in fp32) that lead the CTA's shared memory
    assert shape(4096) == (14, 44)            # two rounds either way: the smaller, faster group
    assert shape(16384) == (16, 26) and shape(65536) == (16, 26)   # SMs stay full: more env-warps, fewer shared rows
    assert shape(1024) == (7, 48) and shape(256) == (2, 48) and shape(100) == (1, 48)   # below one round: all SMs, small groups
    assert shape(2048) == (14, 44)            # exactly one round
    assert shape(4096, dtype=1)[0] in (7, 8) and shape(64, dtype=1) == (1, 48)
    w, r = C.c_int(), C.c_int()
    assert lib.b2h_choose_launch_shape(4096, 148, 8 * 1024, 0, C.byref(w), C.byref(r)) < 0   # scratch does not fit
    assert lib.b2h_choose_launch_shape(0, 148, 232448, 0, C.byref(w), C.byref(r)) < 0


def test_ppo_parameter_layout_matches_python(lib):
    """The flat parameter vector of the PPO update kernels (b2h_ppo_param_layout) is the one policy.py allocates."""
    from mujocoposelearning_b200 import abi
    from mujocoposelearning_b200.policy import param_layout
    assert lib.b2h_sizeof_ppo_config() == C.sizeof(abi.B2HPpoConfig)
    for dims in ((352, 256, 21), (53, 64, 21), (40, 128, 3)):
        offs = (C.c_int64 * 13)()
        n = lib.b2h_ppo_param_layout(*dims, offs)
        shapes, want, total = param_layout(*dims)
        assert list(offs) == want and n == total and all(o % 4 == 0 for o in want)
    assert 317995 <= param_layout(352, 256, 21)[2] < 317995 + 13 * 4      # 317 995 parameters (+ alignment padding)
    cfg = abi.B2HPpoConfig()
    h = C.c_void_p()
    assert lib.b2h_ppo_create(C.byref(cfg), C.byref(h)) == abi.EINVAL      # zero shapes are refused before any device work


def test_policy_parameters_round_trip_through_the_sb3_state_dict():
    """MlpPolicyParams <-> the keys of SB3's ActorCriticPolicy (what PPO.load(...).policy.state_dict() holds): a policy trained
    here loads into the reference's tooling and back; all tensors are views of one flat vector."""
    import torch
    from mujocoposelearning_b200.policy import MlpPolicyParams
    p = MlpPolicyParams(obs_dim=53, act_dim=21, hidden=64, device="cpu", seed=3)
    p.log_std.copy_(torch.linspace(-1, 0, 21))
    sd = p.to_sb3_state_dict()
    assert sd["mlp_extractor.policy_net.0.weight"].shape == (64, 53) and sd["action_net.weight"].shape == (21, 64) and sd["value_net.weight"].shape == (1, 64)
    q = MlpPolicyParams.from_sb3_state_dict(sd, device="cpu")
    assert torch.equal(q.flat, p.flat) and q.offsets == p.offsets
    q.flat.zero_()
    assert all(float(t.abs().sum()) == 0 for t in q.pi + q.vf + [q.log_std])      # views, not copies
