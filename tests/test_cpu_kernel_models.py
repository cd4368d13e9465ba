"""CPU: the GPU-free checks of the tcgen05 kernels that have not run on a GPU yet (lt_attn_tc3_kernel = the "ahead" long-term
attention layout, gp_attn_tc_kernel = the fused DeAOT long-term attention):
  * discrete-event models of their mbarrier / TMEM-buffer / shared-memory-ring protocols (no deadlock, no hazard);
  * an address-level functional model of the tensor-core data path (TMA tiles, UMMA descriptors, TMEM columns, P aliasing),
    calibrated on the arithmetic of the kernel that IS validated on B200, under which both kernels must reproduce attention."""
import importlib.util
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REPO, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_ahead_layout_protocol_model():
    m = _load("lt_ahead_protocol_sim")
    for T in range(0, 9):
        for seed in range(25):
            m.Sim(T, seed * 7919 + T).run()


def test_fused_deaot_kernel_protocol_model():
    m = _load("gp_attn_protocol_sim")
    for T in range(0, 11):
        for seed in range(25):
            m.Sim(T, seed * 104729 + T).run()


def test_tensor_core_layout_model():
    m = _load("tc_layout_model")
    rng = np.random.default_rng(1)
    Q = (rng.standard_normal((200, 32)) * 3).astype(np.float32)
    K = rng.standard_normal((300, 32)).astype(np.float32)
    V = rng.standard_normal((300, 32)).astype(np.float32)
    ref = m.reference(Q, K, V, np.sqrt(32.0))
    assert np.abs(m.model_lt_tile(Q, K, V, ahead=False) - ref).max() < 1e-4      # calibration: the validated kernel
    assert np.abs(m.model_lt_tile(Q, K, V, ahead=True) - ref).max() < 1e-4
    Q = (rng.standard_normal((100, 128)) * 2).astype(np.float32)
    K = rng.standard_normal((200, 128)).astype(np.float32)
    V = rng.standard_normal((200, 256)).astype(np.float32)
    ref = m.reference(Q, K, V, np.sqrt(128.0))
    for vs in range(2):
        assert np.abs(m.model_gp(Q, K, V, vs) - ref[:, vs * 128:(vs + 1) * 128]).max() < 1e-4
