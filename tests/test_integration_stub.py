"""INTEGRATION.md section 2 shows the ctypes binding a reference maintainer would add (`rcognita/_b200.py`).
This test executes that very code block against stand-ins that carry the reference classes' attributes
(System.name / pars / ctrl_bnds / dim_input; CtrlOptPred.mode / critic_struct / ... / w_critic) and checks the
result against the CPU oracle -- so the documented binding is known to work as printed."""
import os
import re
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    src = next(b for b in blocks if "rcognita/_b200.py" in b)
    return src.replace('C.CDLL("librcg_b200.so")', f'C.CDLL("{os.path.join(ROOT, "rcognita_b200", "librcg_b200.so")}")')


def test_stub_struct_layouts_match_the_header():
    """The structs printed in INTEGRATION.md have the sizes of the library's own ctypes mirror (and of rcg.h)."""
    import ctypes as C
    src = stub_source()
    ns = {}
    exec("import ctypes as C\n" + src[src.index("class RcgSystem"):src.index("SYS_ID =")], ns)   # the two struct classes
    from rcognita_b200 import _C
    assert C.sizeof(ns["RcgSystem"]) == C.sizeof(_C.RcgSystem)
    assert C.sizeof(ns["RcgObjective"]) == C.sizeof(_C.RcgObjective)
    assert [f[0] for f in ns["RcgObjective"]._fields_] == [f[0] for f in _C.RcgObjective._fields_]


@pytest.mark.gpu
def test_documented_binding_runs_and_matches_the_oracle():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    import oracle
    ns = {}
    exec(compile(stub_source(), "INTEGRATION.md:rcognita/_b200.py", "exec"), ns)
    bnds = np.array([[-25.0, 25.0], [-5.0, 5.0]])
    system = types.SimpleNamespace(name="3wrobotNI", pars=[], ctrl_bnds=bnds, dim_input=2)
    for mode, cs in (("MPC", "quad-nomix"), ("RQL", "quad-lin")):
        dimc = {"quad-nomix": 5, "quad-lin": 20}[cs]
        ctrl = types.SimpleNamespace(mode=mode, critic_struct=cs, stage_obj_struct="quadratic", Nactor=6, Ncritic=4,
                                     buffer_size=10, gamma=0.95, pred_step_size=0.01, dim_output=3, dim_input=2,
                                     stage_obj_pars=[np.diag([1.0, 10.0, 1.0, 0.0, 0.0])], observation_target=[],
                                     dim_critic=dimc, w_critic=np.linspace(0.1, 2.0, dimc))
        rng = np.random.default_rng(3)
        obs = rng.uniform(-5, 5, size=(17, 3))
        xs = obs + 0.01 * rng.normal(size=obs.shape)
        table = rng.uniform(np.tile(bnds[:, 0], 6), np.tile(bnds[:, 1], 6), size=(40, 12))
        J, am = ns["actor_cost_table"](ctrl, system, obs, xs, table)
        s = oracle.make_sys("3wrobotNI", [], bnds)
        c = oracle.make_ctrl(3, 2, mode=mode, Nactor=6, pred_step_size=0.01, gamma=0.95, critic_struct=cs,
                             R1=[1, 10, 1, 0, 0])
        for e in range(17):
            Jr, ar = oracle.actor_cost_table(c, s, table, obs[e], xs[e], ctrl.w_critic)
            assert np.max(np.abs(J[e] - Jr) / np.abs(Jr)) <= 1e-9
            assert am[e] == int(np.argmin(J[e]))


@pytest.mark.gpu
def test_readme_quick_tour_runs_as_printed():
    """The README's quick-tour block, executed as printed (only the batch and the episode length are shrunk)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    md = open(os.path.join(ROOT, "README.md")).read()
    src = next(b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "ClosedLoopEngine" in b)
    assert "size=(65536, 3)" in src and "t1=10.0" in src
    src = src.replace("size=(65536, 3)", "size=(600, 3)").replace("t1=10.0", "t1=0.3")
    ns = {}
    exec(src, ns)
    res, rows, ckpt = ns["res"], ns["rows"], ns["ckpt"]
    assert res["y"].shape == (600, 3) and np.all(res["status"] == 1) and np.all(res["t"] >= 0.3)
    assert rows.shape[1] == 8 and rows.shape[0] == int(res["nsteps"][7]) // 4 and rows[-1, 0] <= res["t"][7]
    assert ckpt["blob"].numel() > 0 and ckpt["E"] == 600
