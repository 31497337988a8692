"""Organization sharding: the org-sharded round equals the single-rank round.

Two ranks are emulated on ONE GPU, sequentially (no kernels wait on each other): each emulated rank trains and
predicts only its own block of organizations, the exchange is emulated by copying the rows a real all-gather would
deliver, then both run the replicated update. F_t must be IDENTICAL to the single-rank run (same kernels, same
per-organization seeds, same summation order) — bit-for-bit, not just within tolerance.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_emulated_ranks_match_one():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import roundloop, runner, synth
    from dmtcdr_b200.config import make_cfg

    control = "Amazon_user_implicit_ae_0_genre_assist_constant-0.1_constant"
    make_cfg(control, device="cuda", seed=0)
    data = synth.make_rating_data("tiny-Amazon", seed=0)
    torch.manual_seed(0)
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    split = [s.numpy() for s in runner.split_dataset(dataset)]
    mats = {k: (dataset[k].data, dataset[k].target) for k in dataset}
    kw = dict(target_mode="implicit", batch_rows=50, clamp=True, ar=0.1, local_epochs=2, device="cuda:0", seed=3)
    one = roundloop.AssistRounds(mats, split, rank=0, world=1, **kw)
    ranks = [roundloop.AssistRounds(mats, split, rank=r, world=2, **kw) for r in (0, 1)]
    assert sorted(ranks[0].my_orgs + ranks[1].my_orgs) == list(range(len(split)))
    for r in [one] + ranks:
        r.round0()
    for k in ("train", "test"):
        assert torch.equal(one.F[k], ranks[0].F[k]) and torch.equal(one.F[k], ranks[1].F[k])
    for t in (1, 2):
        one.run_round(t)
        for r in ranks:
            r.train_predict(t)
        for r in ranks:
            r.sync()
        # what dist.exchange_outputs (in-place all-gather of equal row blocks) delivers
        for k in ("train", "test"):
            c = ranks[0].chunk
            for src in (0, 1):
                dst = ranks[1 - src]
                dst.state.O_full[k][src * c:(src + 1) * c].copy_(ranks[src].state.O_full[k][src * c:(src + 1) * c])
        for r in ranks:
            r.combine()
        one.sync()
        for k in ("train", "test"):
            assert torch.equal(one.state.O_orgmajor(k), ranks[0].state.O_orgmajor(k))
            assert torch.equal(one.F[k], ranks[0].F[k]) and torch.equal(one.F[k], ranks[1].F[k])
            assert torch.isfinite(one.F[k]).all()
    for r in [one] + ranks:
        r.close()


def test_group_lockstep_matches_per_organization_graphs():
    """dmt_group_train (one launch per step kernel for all organizations) against the per-organization epoch graphs:
    same seeds -> same parameters and predictions up to the summation order of the gradient norm."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import roundloop, runner, synth
    from dmtcdr_b200.config import make_cfg

    control = "Douban_user_explicit_ae_0_genre_assist_constant-0.3_constant"
    make_cfg(control, device="cuda", seed=0)
    data = synth.make_rating_data("tiny-Douban", seed=0)
    torch.manual_seed(0)
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    split = [s.numpy() for s in runner.split_dataset(dataset)]
    mats = {k: (dataset[k].data, dataset[k].target) for k in dataset}
    kw = dict(target_mode="explicit", batch_rows=40, clamp=True, ar=0.3, local_epochs=3, device="cuda:0", seed=5)
    a = roundloop.AssistRounds(mats, split, group=False, **kw)
    b = roundloop.AssistRounds(mats, split, group=True, **kw)
    assert b.group is not None
    for eng in a.eng.values():  # the group launches use the gather decoder: compare like with like
        eng.set_decoder("gather")
    for r in (a, b):
        r.round0()
    for t in (1, 2):
        a.run_round(t)
        b.run_round(t)
        a.sync()
        b.sync()
        for org in a.my_orgs:
            la, lb = a.round_losses[t][org].cpu(), b.round_losses[t][org].cpu()
            assert float((la - lb).abs().max()) <= 1e-5 * float(la.abs().max())
            pa, pb = a.eng[org].params().cpu(), b.eng[org].params().cpu()
            assert float((pa - pb).abs().max()) <= 5e-4 * float(pa.abs().max())
        for k in ("train", "test"):
            assert float((a.F[k] - b.F[k]).abs().max()) <= 1e-4 * float(a.F[k].abs().max())
    a.close()
    b.close()


def test_more_ranks_than_organizations():
    """8 emulated ranks for 4 organizations (ranks 4-7 own nothing, as ranks 6-7 do for 18 organizations on 8 GPUs):
    empty ranks still take part in the exchange and the replicated update; F_t equals the single-rank run bit-for-bit."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import roundloop, runner, synth
    from dmtcdr_b200.config import make_cfg

    control = "Amazon_user_implicit_ae_0_genre_assist_constant-0.1_constant"
    make_cfg(control, device="cuda", seed=0)
    data = synth.make_rating_data("tiny-Amazon", seed=0)
    torch.manual_seed(0)
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    split = [s.numpy() for s in runner.split_dataset(dataset)]
    mats = {k: (dataset[k].data, dataset[k].target) for k in dataset}
    kw = dict(target_mode="implicit", batch_rows=50, clamp=True, ar=0.1, local_epochs=2, device="cuda:0", seed=3)
    one = roundloop.AssistRounds(mats, split, rank=0, world=1, **kw)
    world = 8
    ranks = [roundloop.AssistRounds(mats, split, rank=r, world=world, **kw) for r in range(world)]
    assert sorted(sum((r.my_orgs for r in ranks), [])) == list(range(len(split)))
    assert any(not r.my_orgs for r in ranks)
    for r in [one] + ranks:
        r.round0()
    for t in (1, 2):
        one.run_round(t)
        for r in ranks:
            r.train_predict(t)
        for r in ranks:
            r.sync()
        c = ranks[0].chunk
        for k in ("train", "test"):
            for src in range(world):
                for dst in range(world):
                    if dst != src:
                        ranks[dst].state.O_full[k][src * c:(src + 1) * c].copy_(
                            ranks[src].state.O_full[k][src * c:(src + 1) * c])
        for r in ranks:
            r.combine()
        one.sync()
        for k in ("train", "test"):
            for r in ranks:
                assert torch.equal(one.F[k], r.F[k])
    m0, m7 = ranks[0].evaluate("test"), ranks[7].evaluate("test")
    assert m0 == m7 == one.evaluate("test")
    for r in [one] + ranks:
        r.close()
