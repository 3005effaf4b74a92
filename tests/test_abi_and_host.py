"""CPU-side tests: the C-ABI library loads and exports every symbol include/rcg.h declares,
argument validation and the no-device error path (no compute without a GPU), the sharding
partitioner, shard-invariant synthetic inputs, and the end-of-run collectives over gloo with
world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from rcognita_b200.build import build_library
    return build_library()


def test_header_symbols_are_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "rcg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(rcg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in rcg.h but not exported"
    from rcognita_b200 import _C
    assert sorted(_C.EXPORTS) == declared
    assert L.rcg_version() == 100


def test_struct_layouts_match_header():
    """ctypes mirrors of rcg_system_t / rcg_objective_t / rcg_solver_t / rcg_log_t must have the C sizes."""
    from rcognita_b200 import _C
    src = r'''
    #include <stdio.h>
    #include "rcg.h"
    int main(void) { printf("%zu %zu %zu %zu %zu\n", sizeof(rcg_system_t), sizeof(rcg_objective_t), sizeof(rcg_solver_t), sizeof(rcg_log_t), sizeof(rcg_disturb_t)); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_C.RcgSystem), ctypes.sizeof(_C.RcgObjective), ctypes.sizeof(_C.RcgSolver),
                     ctypes.sizeof(_C.RcgLog), ctypes.sizeof(_C.RcgDisturb)]


def test_dim_helpers_and_descriptors():
    from rcognita_b200 import _C
    assert [_C.lib.rcg_dim_state(i) for i in range(3)] == [3, 5, 2]
    assert [_C.lib.rcg_dim_input(i) for i in range(3)] == [2, 2, 1]
    assert _C.lib.rcg_dim_state(7) == -1
    # controllers.py:1024-1039
    assert [_C.dim_critic(cs, 3, 2) for cs in ("quad-lin", "quadratic", "quad-nomix", "quad-mix")] == [20, 15, 5, 11]
    assert [_C.dim_critic(cs, 5, 2) for cs in ("quad-lin", "quadratic", "quad-nomix", "quad-mix")] == [35, 28, 7, 17]
    assert [_C.dim_critic(cs, 2, 1) for cs in ("quad-lin", "quadratic", "quad-nomix", "quad-mix")] == [9, 6, 3, 5]
    o = _C.make_objective(3, 2, mode="RQL", Nactor=6, gamma=0.9, Ncritic=4, buffer_size=4, R1=[1, 10, 1, 0, 0])
    assert o.Ncritic == 3                                       # min(Ncritic, buffer_size - 1)
    assert o.r_is_diag == 1 and o.gamma_pow[3] == 0.9 ** 3.0
    o = _C.make_objective(3, 2, R1=np.ones((5, 5)))
    assert o.r_is_diag == 0
    with pytest.raises(ValueError):
        _C.make_objective(3, 2, R1=np.ones((4, 4)))
    with pytest.raises(ValueError):
        _C.make_objective(3, 2, Nactor=65)
    s = _C.make_system("3wrobot", [10, 1], [[-300, 300], [-100, 100]])
    assert s.sys_id == 1 and s.has_bnds == 1 and s.hi[1] == 100
    assert _C.make_system("2tank", [1, 2, 3, 4, 5], []).has_bnds == 0


def test_no_device_is_a_loud_error():
    """Without a CUDA device every compute entry point fails with RCG_ENODEV -- there is no
    CPU fallback (and the host wrappers refuse non-CUDA tensors)."""
    import torch
    from rcognita_b200 import _C, ops
    sysd = _C.make_system("3wrobotNI", [], [[-25, 25], [-5, 5]])
    y = torch.zeros((3, 4), dtype=torch.float64)
    a = torch.zeros((2, 4), dtype=torch.float64)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.state_dyn(sysd, y, a)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    assert _C.lib.rcg_device_count() in (0, _C.lib.rcg_device_count()) and _C.lib.rcg_device_count() <= 0
    buf = (ctypes.c_double * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = _C.lib.rcg_state_dyn(ctypes.byref(sysd), 4, p, p, p, None)
    assert rc == -2 and "no CPU fallback" in _C.last_error()
    rc = _C.lib.rcg_state_dyn(ctypes.byref(sysd), 4, None, p, p, None)
    assert rc == -1 and "null" in _C.last_error()
    # round-2 entry points: disturbance lanes, fp32 critic cost
    distd = _C.make_disturb([[1, 1], [0, 0], [0.3, 0.3]], seed=3, env_offset=7)
    assert distd.seed == 3 and distd.env_offset == 7 and distd.tau[1] == 0.3
    rc = _C.lib.rcg_rhs_disturbed(ctypes.byref(sysd), ctypes.byref(distd), 4, p, p, None, None, p, 1, None)
    assert rc == -2 and "no CPU fallback" in _C.last_error()
    assert _C.lib.rcg_rhs_disturbed(ctypes.byref(sysd), None, 4, p, p, None, None, p, 1, None) == -1
    assert _C.lib.rcg_disturb_normals(ctypes.byref(distd), 4, None, 0, p, None) == -2
    sol = _C.make_solver(1.0, 0.005, 1e-3, 1e-5)
    assert _C.lib.rcg_rk45_step_disturbed(ctypes.byref(sysd), None, ctypes.byref(sol), 4, p, p, p, p, p, p, p, None) == -1
    assert "disturbance" in _C.last_error()
    assert _C.lib.rcg_rk45_step_disturbed(ctypes.byref(sysd), ctypes.byref(distd), ctypes.byref(sol), 4, p, p, p, p, p, None, p, None) == -1
    assert "nfev" in _C.last_error()                      # nfev numbers the random draws: required
    obj = _C.make_objective(3, 2, mode="RQL", Nactor=3, R1=[1, 10, 1, 0, 0])
    assert _C.lib.rcg_critic_cost_f32(ctypes.byref(obj), 3, 2, 4, 1, p, p, p, p, p, None) == -2
    assert _C.lib.rcg_last_actor_kernel() in (b"", b"actor_cost_kernel", b"actor_cost_tma_kernel", b"actor_cost_tma_rt_kernel")
    from rcognita_b200.engine import ClosedLoopEngine
    with pytest.raises(RuntimeError, match="CUDA device"):
        ClosedLoopEngine("3wrobotNI", np.zeros((2, 3)), np.zeros((4, 12)))


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under rcognita_b200/ may import or link it."""
    pkg = os.path.join(ROOT, "rcognita_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "rcg_oracle" not in text, f
    out = subprocess.run(["ldd", os.path.join(pkg, "librcg_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_structured_candidates_table():
    from rcognita_b200.controllers import structured_candidates
    for bnds, N, C in (([[-25, 25], [-5, 5]], 6, 256), ([[0, 1]], 8, 64), ([[-300, 300], [-100, 100]], 10, 100), ([[-25, 25], [-5, 5]], 3, 16)):
        b = np.array(bnds, dtype=float)
        m = b.shape[0]
        tab = structured_candidates(bnds, N, C, seed=3)
        assert tab.shape == (C, N * m)
        assert np.all(tab >= np.tile(b[:, 0], N)) and np.all(tab <= np.tile(b[:, 1], N))
        const = tab[np.all(tab.reshape(C, N, m) == tab.reshape(C, N, m)[:, :1, :], axis=(1, 2))]
        assert len(const) >= min(C, 9)                                  # the grid part: constant sequences
        mid = 0.5 * (b[:, 0] + b[:, 1])
        assert any(np.allclose(r[:m], mid) for r in const)              # holding the mid-point is a candidate
        if C >= 3 ** m:
            assert any(np.allclose(r[:m], b[:, 1]) for r in const) and any(np.allclose(r[:m], b[:, 0]) for r in const)
        assert np.array_equal(tab, structured_candidates(bnds, N, C, seed=3))


def test_shard_range_partition():
    from rcognita_b200 import shard
    for E in (1024, 65536, 1048576, 5 * 1024):
        for W in (1, 2, 3, 4, 8):
            if E // 1024 < W:
                continue
            rs = [shard.shard_range(E, r, W) for r in range(W)]
            assert rs[0][0] == 0 and rs[-1][1] == E
            assert all(rs[i][1] == rs[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1024 and all(s % 1024 == 0 for s in sizes)
    with pytest.raises(ValueError):
        shard.shard_range(1000, 0, 2)
    with pytest.raises(ValueError):
        shard.shard_range(2048, 2, 2)


def test_synthetic_inputs_are_shard_invariant():
    import bench_workload as bw
    from rcognita_b200 import shard
    E = 8 * 1024
    full = bw.synthetic_states("3wrobotNI", 0, E, seed=0)
    assert full.shape == (E, 3) and np.all(np.abs(full[:, :2]) <= 10) and np.all(np.abs(full[:, 2]) <= np.pi)
    cfull = bw.synthetic_candidates([[-25, 25], [-5, 5]], 6, 16, seed=1, env_range=(0, E))
    assert cfull.shape == (E, 16, 12)
    assert np.all(np.abs(cfull[..., 0::2]) <= 25) and np.all(np.abs(cfull[..., 1::2]) <= 5)
    for W in (2, 4, 8):
        parts, cparts = [], []
        for r in range(W):
            lo, hi = shard.shard_range(E, r, W)
            parts.append(bw.synthetic_states("3wrobotNI", lo, hi, seed=0))
            cparts.append(bw.synthetic_candidates([[-25, 25], [-5, 5]], 6, 16, seed=1, env_range=(lo, hi)))
        assert np.array_equal(np.concatenate(parts), full)
        assert np.array_equal(np.concatenate(cparts), cfull)
    tab = bw.synthetic_candidates([[0, 1]], 8, 256, seed=1)
    assert tab.shape == (256, 8) and tab.min() >= 0 and tab.max() <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from rcognita_b200 import shard
import bench_workload as bw
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
E = 3 * 1024                                   # uneven: rank 0 owns 2 blocks, rank 1 owns 1
lo, hi = shard.shard_range(E, rank, world)
x0 = bw.synthetic_states("3wrobotNI", lo, hi, seed=0)
local_ret = torch.from_numpy(x0[:, 0] * 2.0 + x0[:, 1])          # stand-in for per-env returns
counts = torch.tensor([hi - lo, 7 * (rank + 1)], dtype=torch.int64)
ret, tot = shard.gather_returns(local_ret, counts)
full = bw.synthetic_states("3wrobotNI", 0, E, seed=0)
assert ret.shape == (E,), ret.shape
assert np.array_equal(ret.numpy(), full[:, 0] * 2.0 + full[:, 1])   # global env order, bit-exact
assert tot.tolist() == [E, 7 * 3], tot.tolist()
# trajectory rings: [capacity, ncols, E_local] per rank -> global environment order on rank 0
cap, ncols = 4, 8
rows = torch.from_numpy(np.arange(lo, hi, dtype=np.float64))[None, None, :] + torch.arange(cap * ncols, dtype=torch.float64).reshape(cap, ncols, 1) * 1e6
cnt = torch.arange(lo, hi, dtype=torch.int32) % 7
R, Cn = shard.gather_trajectories(rows, cnt)
if rank == 0:
    assert R.shape == (cap, ncols, E) and Cn.shape == (E,)
    assert np.array_equal(R[2, 3].numpy(), np.arange(E) + (2 * ncols + 3) * 1e6)
    assert np.array_equal(Cn.numpy(), np.arange(E) % 7)
    e = 2050                                     # owned by rank 1; count 6 > capacity 4 -> wrapped ring
    assert Cn[e] == 6 and np.array_equal(shard.ring_rows(R, Cn, e)[:, 0].numpy(), e + np.array([2, 3, 0, 1]) * ncols * 1e6)
else:
    assert R is None and Cn is None
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gather_returns_gloo_world2(tmp_path):
    """N>1 path on CPU: two gloo ranks shard the batch, gather the per-env returns in global
    order and sum their counters (the only collectives on the path)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (oracle port, all host threads) prints the contract's JSON line."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "1", "--envs", "1024", "--cpu-sample-envs", "1024"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "evals/s"
