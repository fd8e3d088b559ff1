"""Multi-rank halo exchange: world_size 2 and 4 over torch.distributed gloo on the CPU,
with the TEST-ONLY host-emulation build whose transport calls back into this harness.  One fv3jedi_lm
dynamics step (NL, TL, AD) sharded over the ranks must reproduce the single-rank result, and the
distributed dot-product test <M dx, y> = <dx, M^T y> must hold across ranks.  (The NCCL transport of the
product build is exercised by the -m gpu test below with torchrun-style workers when >= 2 GPUs exist.)"""
import os
import sys
import tempfile
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs(nonhydro, two_sided=False):
    """two_sided: both flag structures at the reference's defaults (monotone hord 9 / 12 and kord 8 trajectory, linear increment with
    its own damping and sponge), see test_dyn_core.REF_DEFAULTS"""
    from test_dyn_core import CFG, REF_DEFAULTS, two_sided_params
    from test_fv_dynamics import eta, api_state, ZVIR, RD
    from oracle.cubed_sphere import R
    N, K = 12, (4 if two_sided else 3)
    ak, bk = eta(K, CFG["ptop"])
    f, rng = api_state(N, K, 77, ak, bk, nonhydro)
    act = ["u", "v", "t", "delp", "qv", "ql", "qi", "o3"] + (["w", "delz"] if nonhydro else [])
    for k in act:
        z = np.zeros_like(f[k]); z[..., R(1, N), R(1, N)] = f[k][..., R(1, N), R(1, N)]; f[k] = z
    names = act + ([] if nonhydro else ["w"]) + ["phis"]
    f = {k: f[k] for k in names}
    cfg = dict(CFG); cfg.update(zvir=ZVIR, k_split=1, n_split=2, dt=900.0, hord_tr=2, rdgas=RD, grav=9.80665, p_fac=0.05)
    if two_sided:
        cfg.update(REF_DEFAULTS)
    p = two_sided_params(cfg); p.update(do_vort_damp=1, hydrostatic=0 if nonhydro else 1, nq=4, bdt=900.0)
    if two_sided:
        p["_oracle_cfg"] = cfg          # (make_golden reads it; removed before the parameters reach the library)
    dx = {k: rng.standard_normal(f[k].shape) * (np.abs(f[k]).mean() * 1e-3 + 1e-30) for k in act}
    y = {k + "_n": np.zeros_like(f[k]) for k in act}
    for k in act:
        y[k + "_n"][..., R(1, N), R(1, N)] = rng.standard_normal((6, K, N, N))
    return N, K, ak, bk, f, act, p, dx, y


def _run_all(h, N, K, f, act, p, dx, y, module="step"):
    """NL / TL / AD of a module on handle h (any decomposition); returns global arrays"""
    import fv3lm
    p = {k: (int(v) if isinstance(v, bool) else v) for k, v in p.items() if not k.startswith("_")}
    NX = N + 7
    outs = list(y.keys())
    zero = lambda: np.zeros((6, K, NX, NX))
    res = {}
    traj = {k: h.scatter(f[k]) for k in f}
    for o in outs:
        traj[o] = h.scatter(zero())
    pert = {k: h.scatter(dx[k]) for k in act}
    for o in outs:
        pert[o] = h.scatter(zero())
    h.module_run(module, fv3lm.MODE_TL, traj, pert, params=p)
    for o in outs:
        res["nl." + o] = h.gather(traj[o], zero(), closed=False)
        res["tl." + o] = h.gather(pert[o], zero(), closed=False)
    traj = {k: h.scatter(f[k]) for k in f}
    for o in outs:
        traj[o] = h.scatter(zero())
    pert = {k: h.scatter(np.zeros_like(f[k])) for k in act}
    for o in outs:
        pert[o] = h.scatter_owned(y[o])
    h.module_run(module, fv3lm.MODE_AD, traj, pert, params=p)
    for k in act:
        res["ad." + k] = h.gather_add(pert[k], np.zeros_like(f[k]))
    return res


def _inputs_tracer():
    """tracer_2d with q_split = 0; the Courant numbers that set ksplt(k) = 2 and 3 sit on tiles 5 and 6 only, so with two ranks
    (tiles 1-3 / 4-6) rank 0 gets the sub-step counts from the exchange of the level maxima (mp_reduce_max)"""
    import test_tracer_2d as tt
    from test_dyn_core import CFG
    from oracle.cubed_sphere import R
    N, K = 12, 6
    f, rng = tt.fields(N, K, [0.5, 0.3, 0.6, 0.2, 0.4, 0.7], 5)
    for tile, k, fac in ((4, 1, 4.5), (5, 3, 12.0), (5, 4, 3.0)):
        for n in ("cx", "cy", "mfx", "mfy"):
            f[n][tile, k] *= fac
    act = list(f.keys())
    p = dict(CFG); p.update(hord_tr=2, q_split=0, q_split_max=3)
    dx = {k: rng.standard_normal(f[k].shape) * (np.abs(f[k]).mean() * 1e-3 + 1e-30) for k in act}
    y = {o: np.zeros((6, K, N + 7, N + 7)) for o in ("q0_n", "q1_n")}
    for o in y:
        y[o][..., R(1, N), R(1, N)] = rng.standard_normal((6, K, N, N))
    return N, K, None, None, f, act, p, dx, y


def _inputs_c2l():
    """cubed_to_latlon: the D-grid halo update (mode = 1) crosses the rank boundary; a11 .. a22 are inputs without increments"""
    from common import metrics, rnd
    from oracle.cubed_sphere import R
    N, K = 12, 2
    rng = np.random.default_rng(8)
    M = metrics(N)
    f = dict(u=10.0 * rnd(rng, N, K), v=10.0 * rnd(rng, N, K))
    for n in ("a11", "a12", "a21", "a22"):
        f[n] = np.ascontiguousarray(M[n][:, None])
    act = ["u", "v"]
    dx = {k: rng.standard_normal(f[k].shape) for k in act}
    y = {o: np.zeros((6, K, N + 7, N + 7)) for o in ("ua", "va")}
    for o in y:
        y[o][..., R(1, N), R(1, N)] = rng.standard_normal((6, K, N, N))
    return N, K, None, None, f, act, {}, dx, y


_MODULE_INPUTS = {"tracer_2d": _inputs_tracer, "c2l_ord4": _inputs_c2l}


def _worker(rank, world, port, nonhydro, outdir, two_sided=False):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "fv3-jedi-linearmodel_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import fv3lm
    from common import metrics
    module = nonhydro if nonhydro in _MODULE_INPUTS else "step"
    N, K, ak, bk, f, act, p, dx, y = _MODULE_INPUTS[module]() if module != "step" else _inputs(nonhydro, two_sided)
    cfg = fv3lm.default_config(N, K, rank=rank, nranks=world)
    h = fv3lm.FV3LM(cfg, ak, bk, emu=True)
    h.set_metrics(metrics(N))

    def exchange(peers, sends, recvs):
        reqs = []
        rt = [torch.empty(len(r), dtype=torch.float64) for r in recvs]
        for pr, r in zip(peers, rt):
            if r.numel():
                reqs.append(dist.irecv(r, src=pr))
        for pr, s in zip(peers, sends):
            if len(s):
                reqs.append(dist.isend(torch.from_numpy(np.array(s, copy=True)), dst=pr))
        for q in reqs:
            q.wait()
        for r, dst in zip(rt, recvs):
            if r.numel():
                dst[...] = r.numpy()
    h.comm_set_callback(exchange)
    res = _run_all(h, N, K, f, act, p, dx, y, module)
    nex, nbytes = h.comm_stats()
    assert nex > 0 and nbytes > 0
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


def _check(world, nonhydro, two_sided=False):
    sys.path.insert(0, HERE)
    import fv3lm
    from common import metrics
    module = nonhydro if nonhydro in _MODULE_INPUTS else "step"
    N, K, ak, bk, f, act, p, dx, y = _MODULE_INPUTS[module]() if module != "step" else _inputs(nonhydro, two_sided)
    h = fv3lm.FV3LM(fv3lm.default_config(N, K), ak, bk, emu=True)
    h.set_metrics(metrics(N))
    ref = _run_all(h, N, K, f, act, p, dx, y, module)
    port = 29600 + (os.getpid() % 200)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, nonhydro, d, two_sided), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
        tot = {k: sum(pt[k] for pt in parts) for k in ref}    # every rank wrote only its own cells (zeros elsewhere)
    for k in ref:
        den = max(np.abs(ref[k]).max(), 1e-300)
        e = np.abs(tot[k] - ref[k]).max() / den
        assert e < 1e-11, (k, e)
    # distributed dot-product test
    lhs = sum((tot["tl." + o] * y[o]).sum() for o in y)
    rhs = sum((dx[k] * tot["ad." + k]).sum() for k in act)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs)), (lhs, rhs)


@pytest.mark.parametrize("world,nonhydro", [(2, False), (2, True), (4, True)])
def test_multirank_step_gloo(world, nonhydro):
    _check(world, nonhydro)


def test_multirank_step_two_sided_gloo():
    """two-sided mode (detached views, spliced chains, monotone trajectory schemes) sharded over two ranks"""
    _check(2, True, two_sided=True)


@pytest.mark.gpu
def test_multirank_step_nccl_gpu():
    """product build, NCCL transport: needs >= 2 GPUs (skipped on a single-GPU box)"""
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(HERE)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "multigpu_check.py"), "--nonhydro"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_multirank_tracer_sub_steps_gloo():
    """q_split = 0: the per-level Courant maxima are exchanged between the ranks before the sub-step counts are fixed"""
    _check(2, "tracer_2d")
    import oracle.fv_dynamics as ofv, test_tracer_2d as tt
    from oracle.dyn_core import halo_of
    N, K, _, _, f, act, p, dx, y = _inputs_tracer()
    from common import ograd
    halo, _ = halo_of(N)
    t = {k: torch.from_numpy(v) for k, v in f.items()}
    ofv.tracer_2d([halo.scalar(t["q0"]), halo.scalar(t["q1"])], t["dp1"], t["mfx"], t["mfy"], t["cx"], t["cy"], ograd(N), 2, q_split=0, halo=halo)
    assert ofv.tracer_2d.last_nsplt == 3          # the inputs do need the sub-steps


def test_multirank_c2l_gloo():
    """cubed_to_latlon over two ranks: the halo update of the D-grid winds is a real exchange"""
    _check(2, "c2l_ord4")
