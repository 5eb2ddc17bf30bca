"""CPU suite, part 4: the N>1 host logic under torch.distributed (gloo, world_size 2, 127.0.0.1).

The data path has one exchange step only in consensus mode (estimator.py:226-231: per-position sum of the reads'
log-likelihood chunks).  Here two CPU ranks each own a shard of the reads (shard_reads), plan the overlap groups
with one all_gather_object (plan_groups_host), add their chunks into the concatenated accumulator at the planned
rows, all-reduce it, and must reproduce the single-process grouping, coverage and sums of the reference algorithm
(the oracle's estimate_probabilities keeps the reference's order; a different summation order is allowed 1e-12)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _make_chunks(seed=3):
    rng = np.random.default_rng(seed)
    # three overlap groups, a touching pair (400..450 | 450..520: '>=' opens a new group) and a single
    intervals = [(0, 80), (40, 130), (120, 200), (60, 100), (300, 360), (330, 400), (400, 450), (450, 520),
                 (470, 540), (700, 760), (10, 50)]
    values = [rng.normal(-2, 1, size=(e - s, 4)) for s, e in intervals]
    return intervals, values


def _single_process(intervals, values):
    from nadavca_b200.estimator import group_intervals
    groups = group_intervals(intervals)
    out = []
    for start, end, members in groups:
        acc = np.zeros((end - start, 4))
        cov = np.zeros(end - start, dtype=np.int64)
        for i in sorted(members, key=lambda j: intervals[j]):
            s, e = intervals[i]
            acc[s - start:e - start] += values[i]
            cov[s - start:e - start] += 1
        out.append((start, end, acc, cov))
    return out


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nadavca_b200.estimator import plan_groups_host, shard_reads
        intervals, values = _make_chunks()
        work = [e - s for s, e in intervals]
        mine = shard_reads(work, world)[rank]
        local_iv = [intervals[i] for i in mine]
        groups, group_off, dest = plan_groups_host(local_iv, independent=False, process_group=dist.group.WORLD)
        total = int(group_off[-1])
        acc = torch.zeros((total, 4), dtype=torch.float64)
        cov = torch.zeros(total, dtype=torch.int32)
        for d, i in zip(dest, mine):
            n = intervals[i][1] - intervals[i][0]
            acc[d:d + n] += torch.from_numpy(values[i])
            cov[d:d + n] += 1
        dist.all_reduce(acc)
        dist.all_reduce(cov)
        # independent mode needs no exchange: planning must not communicate and keeps local order
        g2, off2, dest2 = plan_groups_host(local_iv, independent=True, process_group=dist.group.WORLD)
        assert len(g2) == len(local_iv) and dest2.tolist() == off2[:-1].tolist()
        np.savez(os.path.join(result_dir, 'rank%d.npz' % rank), acc=acc.numpy(), cov=cov.numpy(),
                 group_off=group_off, ranges=np.array([(g[0], g[1]) for g in groups]), mine=np.array(mine))
    finally:
        dist.destroy_process_group()


def test_consensus_reduce_two_ranks_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    intervals, values = _make_chunks()
    want = _single_process(intervals, values)
    res = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(world)]
    # every read is owned by exactly one rank
    assert sorted(np.concatenate([r['mine'] for r in res]).tolist()) == list(range(len(intervals)))
    for r in res:
        assert r['ranges'].tolist() == [[s, e] for s, e, _, _ in want]
        for gi, (s, e, acc, cov) in enumerate(want):
            a, b = r['group_off'][gi], r['group_off'][gi + 1]
            assert b - a == e - s
            np.testing.assert_allclose(r['acc'][a:b], acc, rtol=1e-12, atol=1e-12)
            assert np.array_equal(r['cov'][a:b], cov)
    assert np.array_equal(res[0]['acc'], res[1]['acc'])
    # the touching pair opened a new group
    ranges = res[0]['ranges'].tolist()
    assert any(ranges[i][1] == ranges[i + 1][0] for i in range(len(ranges) - 1))


def _bench_rank_worker(rank, world, port, result_dir):
    """bench.py's weak-scaling partition: rank r simulates reads [r*R, (r+1)*R) -- disjoint and deterministic."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import bench
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        assert bench.reduce_over_ranks(t.item(), 'max', None) == float(world)
        assert bench.reduce_over_ranks(t.item(), 'sum', None) == world * (world + 1) / 2
        first = bench.first_read_index(rank, 5)
        np.save(os.path.join(result_dir, 'first%d.npy' % rank), np.array([first]))
    finally:
        dist.destroy_process_group()


def test_bench_rank_helpers_gloo(tmp_path):
    world = 2
    mp.spawn(_bench_rank_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    firsts = [int(np.load(os.path.join(str(tmp_path), 'first%d.npy' % r))[0]) for r in range(world)]
    assert firsts == [0, 5]


def _median_worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nadavca_b200.read import Read, distributed_median
        rng = np.random.default_rng(99)
        results = []
        for n_total in (7, 10, 1, 2001, 4000):
            pooled = np.round(rng.normal(90, 15, size=n_total) * (4 if n_total > 100 else 1)) / 4.0  # with ties
            pooled[::3] *= -1
            mine = pooled[rank::world]
            results.append((distributed_median(mine, dist.group.WORLD), float(np.median(pooled))))
        # Read.normalize_reads over shards == over the pooled reads
        raws = [rng.normal(90, 15, size=int(rng.integers(50, 200))) for _ in range(9)]
        reads = [Read.from_arrays(r, 'ACGT', {0: 0}) for r in raws]
        Read.normalize_reads(reads[rank::world], dist.group.WORLD)
        pooled_reads = [Read.from_arrays(r, 'ACGT', {0: 0}) for r in raws]
        Read.normalize_reads(pooled_reads)
        same = all(np.array_equal(a.normalized_signal, b.normalized_signal)
                   for a, b in zip(reads[rank::world], pooled_reads[rank::world]))
        np.save(os.path.join(result_dir, 'median%d.npy' % rank), np.array(results + [(float(same), 1.0)]))
    finally:
        dist.destroy_process_group()


def test_distributed_median_is_exact_gloo(tmp_path):
    world = 2
    mp.spawn(_median_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = np.load(os.path.join(str(tmp_path), 'median%d.npy' % r))
        assert np.array_equal(res[:, 0], res[:, 1]), res


def _numpy_posterior_rows(ref_codes, group_off, k, prior):
    """posterior_rows callback for consensus_exchange on CPU tensors: _compute_posterior (estimator.py:123-156) for
    the rows [row_lo, row_hi), reading ONLY the local rows it is given (slice + halo)."""
    def fn(local, base_row, row_lo, row_hi, out_rows):
        loc = local.numpy()
        out = out_rows.numpy()
        for g in range(row_lo, row_hi):
            gi = int(np.searchsorted(group_off, g, side='right')) - 1
            gs, ge = int(group_off[gi]), int(group_off[gi + 1])
            i, L = g - gs, ge - gs
            cs, ce = max(0, i - k + 1), min(i + k, L)
            win = loc[gs + cs - base_row:gs + ce - base_row, :4]
            mx = win.max()
            c = 3
            snp = 1 / ((1 - prior) / (prior / c) + (1 - (ce - cs - 1)) * c)
            nonsnp = 1 - snp * c
            pr = np.zeros(4)
            for j in range(4):
                pr[j] = np.exp(loc[g - base_row, j] - mx) * (nonsnp if j == ref_codes[g] else snp)
                if j == ref_codes[g]:
                    for i2 in range(cs, ce):
                        if i2 == i:
                            continue
                        for j2 in range(4):
                            if j2 != ref_codes[gs + i2]:
                                pr[j] += np.exp(loc[gs + i2 - base_row, j2] - mx) * snp
            out[g - row_lo, :4] = pr / sum(pr)
            out[g - row_lo, 4] = loc[g - base_row, 4]
    return fn


def _exchange_worker(rank, world, port, result_dir, collective):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nadavca_b200.estimator import consensus_exchange, plan_groups_host, shard_reads, slice_geometry
        intervals, values = _make_chunks()
        mine = shard_reads([e - s for s, e in intervals], world)[rank]
        groups, group_off, dest = plan_groups_host([intervals[i] for i in mine], False, dist.group.WORLD)
        total = int(group_off[-1])
        slice_rows, total_pad = slice_geometry(total, world)
        rows = torch.zeros((total_pad, 5), dtype=torch.float64)
        for d, i in zip(dest, mine):
            n = intervals[i][1] - intervals[i][0]
            rows[d:d + n, :4] += torch.from_numpy(values[i])
            rows[d:d + n, 4] += 1
        ref_codes = np.random.default_rng(8).integers(0, 4, size=total)
        k, prior = 6, 0.001
        out = consensus_exchange(rows, total, k - 1, dist.group.WORLD, _numpy_posterior_rows(ref_codes, group_off, k, prior),
                                 collective)
        np.savez(os.path.join(result_dir, 'x%d.npz' % rank), out=out.numpy()[:total], group_off=group_off)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('collective,world', [('reduce_scatter', 2), ('allreduce', 2), ('reduce_scatter', 5)])
def test_consensus_exchange_slices_equal_single_process_gloo(tmp_path, collective, world):
    """Reduce-scatter by genome slice + halo rows + posterior on the owned slice + all-gather == the posterior of the
    summed chunks computed in one process (and == the all-reduce variant), on two and on five gloo ranks (five: slices
    that do not divide the rows, halos that cross group boundaries)."""
    from nadavca_b200.estimator import consensus_exchange
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path), collective), nprocs=world, join=True)
    intervals, values = _make_chunks()
    want = _single_process(intervals, values)
    total = sum(e - s for s, e, _, _ in want)
    group_off = np.concatenate([[0], np.cumsum([e - s for s, e, _, _ in want])])
    rows = torch.zeros((total, 5), dtype=torch.float64)
    for gi, (s, e, acc, cov) in enumerate(want):
        rows[group_off[gi]:group_off[gi + 1], :4] = torch.from_numpy(acc)
        rows[group_off[gi]:group_off[gi + 1], 4] = torch.from_numpy(cov.astype(np.float64))
    ref_codes = np.random.default_rng(8).integers(0, 4, size=total)
    single = consensus_exchange(rows, total, 5, None, _numpy_posterior_rows(ref_codes, group_off, 6, 0.001)).numpy()
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), 'x%d.npz' % r))
        assert np.array_equal(got['group_off'], group_off)
        np.testing.assert_allclose(got['out'][:, :4], single[:, :4], rtol=1e-12, atol=1e-300)
        assert np.array_equal(got['out'][:, 4], single[:, 4])
        np.testing.assert_allclose(got['out'][:, :4].sum(axis=1), 1.0, rtol=1e-12)
