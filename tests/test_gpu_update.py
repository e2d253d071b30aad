"""The remaining per-iteration updates on the device (csrc/update.cuh) and whole
runs (sweep + updates) against the CPU oracle (oracle/ppmx_chain.c) under shared Philox streams.

Integer state (labels, counts, K, grid index, acceptance counters) must be identical; continuous
state agrees to 1e-9 relative (device and libm transcendentals differ in the last ulps; nothing
amplifies them unless a Metropolis / allocation decision flips, which has probability ~1e-13 per
decision)."""
import numpy as np
import pytest

import oracle_binding as ob
from cases import build_case

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from treatppmx_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _compare(g, o, dims, what):
    assert (g.nclu == o.nclu).all() and (g.labels == o.labels).all() and (g.idxs == o.idxs).all(), what
    for t in range(dims.T):
        K = int(o.nclu[t])
        assert (g.nj[t, :K] == o.nj[t, :K]).all(), what
        np.testing.assert_allclose(g.eta_star[:, :K, t], o.eta_star[:, :K, t], rtol=RTOL, atol=1e-12, err_msg=what)
    for f in ("loggamma", "JJ", "ss", "beta", "sigma", "kappa", "Theta", "Sigma", "lambda_", "tau", "sigma_beta"):
        np.testing.assert_allclose(getattr(g, f), getattr(o, f), rtol=RTOL, atol=1e-12, err_msg=f"{what}: {f}")


CASES_UPD = ["default_ngg", "dp_T3", "Q3_D3", "nnig", "calib2", "categorical", "ppm_only", "noreuse"]


@pytest.mark.parametrize("name", CASES_UPD)
def test_update_step_matches_oracle(ctx, name):
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case(name, n_datasets=2)
    nchains = 3
    cds = [c % 2 for c in range(nchains)]
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    b.init(19, labels, nclu, betas[0])
    pbs = [ob.Problem(model, dims, shared, d) for d in dsets]
    sts = [pbs[cds[c]].init_state(19, c, labels, nclu, betas[0]) for c in range(nchains)]
    accs = [pbs[cds[c]].new_acc() for c in range(nchains)]
    for l in range(6):
        b.sweep(19, l, 1)
        b.update(19, l)
        for c in range(nchains):
            pbs[cds[c]].sweep(sts[c], 19, c, l, 1)
            pbs[cds[c]].update_iter(sts[c], 19, c, l, accs[c])
            _compare(b.get_state(c), sts[c], dims, f"{name} chain {c} iter {l}")
    for c in range(nchains):
        ea, ba = b.get_acceptance(c)
        assert (ea == accs[c][0]).all() and (ba[: max(dims.Q, 1)] == accs[c][1]).all()


@pytest.mark.parametrize("name", ["default_ngg", "dp_T3"])
def test_run_lockstep_with_oracle(ctx, name):
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case(name)
    nchains, niter = 4, 40
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0] * nchains)
    b.init(23, labels, nclu, betas[0])
    b.run(23, 0, niter)
    pb = ob.Problem(model, dims, shared, dsets[0])
    for c in range(nchains):
        st = pb.init_state(23, c, labels, nclu, betas[0])
        pb.iterate(st, 23, c, 0, niter)
        _compare(b.get_state(c), st, dims, f"{name} chain {c} after {niter} iterations")


@pytest.mark.parametrize("name,groups", [("default_ngg", 3), ("dp_T3", 2), ("nnig", 4)])
def test_run_in_chain_groups_matches_oracle_and_one_group(ctx, name, groups):
    """tppmx_batch_run with the chains split into groups on their own streams (TPPMX_OPT_CHAIN_GROUPS): chains are
    independent, so every chain must end exactly where it does in one group -- and where the oracle does; the
    per-dataset summaries (shared accumulators) must agree too."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case(name, n_datasets=2)
    nchains, niter = 7, 24
    cds = [c % 2 for c in range(nchains)]
    outs = []
    for g in (1, groups):
        b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
        b.init(29, labels, nclu, betas[0])
        b.set_chain_groups(g)
        b.run(29, 0, niter, burn=8, thin=2, psm=True, predictive=dims.npred > 0)
        assert b.last_chain_groups() == g
        outs.append(([b.get_state(c) for c in range(nchains)], b.summaries([1.0] * dims.D)))
        b.close()
    pbs = [ob.Problem(model, dims, shared, d) for d in dsets]
    for c in range(nchains):
        st = pbs[cds[c]].init_state(29, c, labels, nclu, betas[0])
        pbs[cds[c]].iterate(st, 29, c, 0, niter)
        _compare(outs[1][0][c], st, dims, f"{name} chain {c}, {groups} groups vs oracle")
        _compare(outs[1][0][c], outs[0][0][c], dims, f"{name} chain {c}, {groups} groups vs one group")
    s1, sg = outs[0][1], outs[1][1]
    assert np.array_equal(s1["psm"], sg["psm"])
    if dims.npred > 0:
        np.testing.assert_allclose(sg["pi_mean"], s1["pi_mean"], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("over", [dict(hsp=0), dict(upd_hier=0), dict(noprog=1, hsp=0, upd_hier=0)],
                         ids=["hsp0", "upd_hier0", "noprog1-hsp0-upd_hier0"])
def test_run_lockstep_with_the_update_switches_off(ctx, over):
    """The switches of the update part (no horseshoe / frozen hierarchy / no prognostic coefficients):
    tppmx_batch_run (horseshoe on its own stream when on) and single tppmx_batch_update calls
    (horseshoe inside the gamma-draw kernel) against the oracle."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg")
    for k, v in over.items():
        setattr(model, k, v)
    nchains, niter = 3, 12
    pb = ob.Problem(model, dims, shared, dsets[0])
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0] * nchains)
    b.init(29, labels, nclu, betas[0])
    b.run(29, 0, niter)
    for l in range(niter, niter + 3):
        b.sweep(29, l, 1)
        b.update(29, l)
    for c in range(nchains):
        st = pb.init_state(29, c, labels, nclu, betas[0])
        pb.iterate(st, 29, c, 0, niter + 3)
        _compare(b.get_state(c), st, dims, f"{over} chain {c}")


def test_run_accumulates_on_saved_iterations(ctx):
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg", n_datasets=2)
    cds = [0, 0, 1, 1]
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    b.init(2, labels, nclu, betas[0])
    b.reset_summaries()
    b.run(2, 0, 30, burn=10, thin=4, psm=True, predictive=True)   # saved: l in {11,15,19,23,27}
    out = b.summaries([0.0, 40.0, 100.0])
    n_a = np.bincount(dsets[0].treat, minlength=dims.T)
    for ds in range(2):
        for t in range(dims.T):
            d = np.diag(out["psm"][ds, t])[: n_a[t]]
            assert np.allclose(d, 1.0)
            assert np.allclose(out["psm"][ds, t] * 10, np.round(out["psm"][ds, t] * 10))  # 5 saves x 2 chains
    np.testing.assert_allclose(out["pi_mean"].sum(axis=2), 1.0, atol=1e-12)


def test_posterior_agrees_with_oracle_chains(ctx):
    """L3: different seeds, same posterior — the number of clusters per arm from 96 device chains
    against 6 oracle chains, within Monte-Carlo error."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg")
    burn, keep = 150, 150
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0] * 96)
    b.init(1001, labels, nclu, betas[0])
    b.run(1001, 0, burn)
    Kdev = []
    for s in range(0, keep, 5):
        b.run(1001, burn + s, 5)
        Kdev.append(b.get_partitions()[1].astype(float))
    Kdev = np.stack(Kdev)                     # [saves, chains, T]
    pb = ob.Problem(model, dims, shared, dsets[0])
    Kor = []
    for c in range(6):
        st = pb.init_state(77, c, labels, nclu, betas[0])
        pb.iterate(st, 77, c, 0, burn)
        for s in range(0, keep, 5):
            pb.iterate(st, 77, c, burn + s, 5)
            Kor.append(st.nclu.astype(float).copy())
    Kor = np.stack(Kor).reshape(6, -1, dims.T)
    m_dev, m_or = Kdev.mean(axis=(0, 1)), Kor.mean(axis=(0, 1))
    # between-chain standard error of the oracle mean (chains are independent)
    se = Kor.mean(axis=1).std(axis=0, ddof=1) / np.sqrt(6) + Kdev.mean(axis=0).std(axis=0, ddof=1) / np.sqrt(96)
    assert (np.abs(m_dev - m_or) < 4 * se + 0.15).all(), (m_dev, m_or, se)


def test_posterior_summaries_agree_with_oracle_chains(ctx):
    """L3 on the summaries a user reads: co-clustering matrix and posterior-mean predictive
    probabilities accumulated on the device against oracle chains with different seeds.  Every
    device chain is mapped to its own copy of the dataset, so the per-dataset device summaries
    are per-chain values and both sides give a between-chain Monte-Carlo standard error:
    two-sample z-scores must look like noise."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg")
    burn, keep, thin = 200, 300, 3
    nch, nor = 48, 10
    b = Batch(ctx, model, dims, shared, [dsets[0]] * nch, chain_dataset=list(range(nch)))
    b.init(5005, labels, nclu, betas[0])
    b.reset_summaries()
    b.run(5005, 0, burn + keep, burn=burn, thin=thin, psm=True, predictive=True)
    w = np.array([0.0, 40.0, 100.0])
    dev = b.summaries(w)
    pb = ob.Problem(model, dims, shared, dsets[0])
    n_a = np.bincount(dsets[0].treat, minlength=dims.T)
    psm_c = np.zeros((nor, dims.T, dims.ntmax, dims.ntmax))
    pi_c = np.zeros((nor, dims.npred, dims.D, dims.T))
    for c in range(nor):
        st = pb.init_state(909, c, labels, nclu, betas[0])
        pb.iterate(st, 909, c, 0, burn)
        ns = 0
        for l in range(burn, burn + keep):
            pb.iterate(st, 909, c, l, 1)
            if (l + 1) % thin == 0:
                for t in range(dims.T):
                    lab = st.labels[t, : n_a[t]]
                    psm_c[c, t, : n_a[t], : n_a[t]] += lab[:, None] == lab[None, :]
                pi_c[c] += pb.predict(st, 909, c, l)[0]
                ns += 1
        psm_c[c] /= ns
        pi_c[c] /= ns

    def zscores(d, o):   # d: [nch, ...] device per chain, o: [nor, ...] oracle per chain
        se = np.sqrt(d.var(0, ddof=1) / d.shape[0] + o.var(0, ddof=1) / o.shape[0])
        return (d.mean(0) - o.mean(0)) / np.maximum(se, 1e-3)

    z_pi = zscores(dev["pi_mean"], pi_c)
    assert np.abs(z_pi).max() < 5.0 and np.sqrt(np.mean(z_pi ** 2)) < 1.6, (np.abs(z_pi).max(), np.sqrt(np.mean(z_pi ** 2)))
    for t in range(dims.T):
        iu = np.triu_indices(n_a[t], 1)
        d, o = dev["psm"][:, t][:, iu[0], iu[1]], psm_c[:, t][:, iu[0], iu[1]]
        z = zscores(d, o)
        assert np.abs(z).max() < 5.5 and np.sqrt(np.mean(z ** 2)) < 1.6, (t, np.abs(z).max(), np.sqrt(np.mean(z ** 2)))
        assert np.corrcoef(d.mean(0), o.mean(0))[0, 1] > 0.95
    # treatment choice from the pooled posterior means: identical wherever the utility gap is clear
    util_d = np.einsum("k,pkt->tp", w, dev["pi_mean"].mean(0))
    util_o = np.einsum("k,pkt->tp", w, pi_c.mean(0))
    gap = np.sort(util_o, axis=0)[-1] - np.sort(util_o, axis=0)[-2]
    clear = gap > 8.0
    assert (util_d.argmax(0)[clear] == util_o.argmax(0)[clear]).all()


def test_run_stops_at_the_interrupt_flag():
    """tppmx_set_interrupt_flag: a raised flag stops tppmx_batch_run at the next check with
    TPPMX_ERR_INTERRUPTED and leaves the batch in the state after a completed iteration - the same
    state an uninterrupted run of that many iterations reaches."""
    from treatppmx_b200 import _lib
    from treatppmx_b200.engine import Batch, Context
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg")
    ctx = Context(0)
    flag = np.zeros(1, dtype=np.int32)
    ctx.set_interrupt_flag(flag, check_every=5)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 0])
    b.init(3, labels, nclu, betas[0])
    b.run(3, 0, 12)                      # flag down: runs through
    flag[0] = 1
    with pytest.raises(_lib.TppmxError) as ei:
        b.run(3, 12, 40)
    assert ei.value.code == 5 and "iteration 16" in str(ei.value)   # first check after 5 iterations: 12..16 done
    got = b.get_state(0)
    ctx.set_interrupt_flag(None)
    b2 = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 0])
    b2.init(3, labels, nclu, betas[0])
    b2.run(3, 0, 17)
    want = b2.get_state(0)
    assert (got.labels == want.labels).all() and (got.nclu == want.nclu).all()
    np.testing.assert_array_equal(got.beta, want.beta)
    np.testing.assert_array_equal(got.JJ, want.JJ)
    b.close()
    b2.close()
    ctx.close()
