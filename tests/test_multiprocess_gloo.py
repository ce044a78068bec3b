"""CPU, world_size 2, gloo: the N>1 host logic of the path (utterance sharding + gradient all-reduce).
The model is the CPU oracle (the CUDA kernels need a GPU); what is under test is dcasr_b200.distributed."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import REPO, PKG_DIR


def _worker(rank, world, port, out):
    sys.path.insert(0, REPO); sys.path.insert(0, PKG_DIR); sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _util import fill_weights
    from dcasr_b200.distributed import GradAllReducer, shard_indices
    from oracle.encoder_ref import MambaStackRef
    torch.manual_seed(0)
    st = MambaStackRef(1, 64)
    fill_weights(st, 3)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(6, 20, 64, generator=g)                     # 6 utterances in the global batch
    idx = shard_indices(6, rank, world)
    red = GradAllReducer(st.parameters(), bucket_mb=0.05, overlap=(rank >= 0))   # several buckets, hook-driven
    st.zero_grad(set_to_none=True)
    st(X[idx]).pow(2).sum().backward()                          # hooks launch each bucket as it completes
    red()
    st.zero_grad(set_to_none=True)
    st(X[idx]).pow(2).sum().backward()                          # a second step must reuse the buckets cleanly
    red()
    torch.save({k: p.grad.clone() for k, p in st.named_parameters()}, os.path.join(out, f"g{rank}.pt"))
    # gradient accumulation (reference trainer accum_grad = 2): the first micro-batch under no_sync(), as with DDP
    st.zero_grad(set_to_none=True)
    half = len(idx) // 2 or 1
    with red.no_sync():
        st(X[idx[:half]]).pow(2).sum().backward()
    st(X[idx[half:]] if len(idx) > half else X[idx]).pow(2).sum().backward()
    red()
    torch.save({k: p.grad.clone() for k, p in st.named_parameters()}, os.path.join(out, f"acc{rank}.pt"))
    # a second backward WITHOUT no_sync() while buckets are in flight must raise, not publish partial sums
    st.zero_grad(set_to_none=True)
    st(X[idx]).pow(2).sum().backward()
    try:
        st(X[idx]).pow(2).sum().backward()
        err = ""
    except RuntimeError as e:
        err = str(e)
    red()                                                        # drains the collectives launched by the first pass
    torch.save(err, os.path.join(out, f"err{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_equal_mean_of_shards(tmp_path):
    world, port = 2, 29531
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k                     # all ranks hold the same reduced gradient
    # single-process truth: mean over ranks of the per-shard gradients
    from _util import fill_weights
    from dcasr_b200.distributed import shard_indices
    from oracle.encoder_ref import MambaStackRef
    X = torch.randn(6, 20, 64, generator=torch.Generator().manual_seed(5))
    acc = None
    for r in range(world):
        st = MambaStackRef(1, 64)
        fill_weights(st, 3)
        st(X[shard_indices(6, r, world)]).pow(2).sum().backward()
        gr = {k: p.grad for k, p in st.named_parameters()}
        acc = gr if acc is None else {k: acc[k] + gr[k] for k in gr}
    for k in g0:
        assert torch.allclose(g0[k], acc[k] / world, rtol=1e-5, atol=1e-6), k


def test_accumulated_micro_batches_and_double_backward_guard(tmp_path):
    """accum_grad = 2 through no_sync() gives the mean over ranks of the SUM over micro-batches (what DDP gives), and a
    second un-guarded backward is refused (it used to all-reduce a partial sum silently)."""
    world, port = 2, 29533
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a0, a1 = torch.load(tmp_path / "acc0.pt"), torch.load(tmp_path / "acc1.pt")
    g0 = torch.load(tmp_path / "g0.pt")
    for k in a0:
        assert torch.equal(a0[k], a1[k]), k
        # x -> sum of squares is additive over utterances: two micro-batches that partition the shard give the shard's gradient
        assert torch.allclose(a0[k], g0[k], rtol=1e-4, atol=1e-5), k
    for r in range(world):
        assert "no_sync" in torch.load(tmp_path / f"err{r}.pt")


def test_shard_indices_cover_and_balance():
    from dcasr_b200.distributed import shard_indices
    for n, w in ((40, 8), (41, 8), (7, 2), (3, 4)):
        parts = [shard_indices(n, r, w) for r in range(w)]
        assert len({len(p) for p in parts}) == 1
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n - n % w))


def _flag_worker(rank, world, port, out):
    sys.path.insert(0, REPO); sys.path.insert(0, PKG_DIR)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dcasr_b200.trainer_sync import any_rank_flag, patch_trainer

    class T:                                            # stands in for the reference Trainer (world_size attribute only)
        world_size = world
    patch_trainer(T)
    tr = T()
    res = [any_rank_flag(False, world), any_rank_flag(rank == 1, world), any_rank_flag(rank == 0, world), any_rank_flag(True, world),
           tr._any_rank_oom(rank == 1), tr._any_rank_oom(False)]
    torch.save(res, os.path.join(out, f"f{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_host_side_group_flag_matches_reference_semantics(tmp_path):
    """Trainer._any_rank_oom (reference training/trainer.py:200-208): MAX over ranks, one matched collective per call --
    here over a host-side group, so the CUDA stream is never read back."""
    world, port = 2, 29547
    mp.spawn(_flag_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert torch.load(tmp_path / f"f{r}.pt") == [False, True, True, True, True, False]
    from dcasr_b200.trainer_sync import any_rank_flag
    assert any_rank_flag(True, 1) is True and any_rank_flag(False, 1) is False      # one rank: identity, no process group
