"""bench.py --config N for the BASELINE.json configs other than the headline one (configs[1] and the consensus block
of configs[2] live in bench.py itself).  One JSON line per run, same keys as the bench contract where they apply:

  --config 0   nadavca align: 100 synthetic reads (~2 kb) against a 50 kb reference, default config.  Step = the
               device work of one alignment round of align_signal for the whole batch: refine_alignment(transitions)
               + per-event means (align_signal.py:52-70); `api` = the public align_signal() call (two alignment
               rounds + renormalisation, host work included).
  --config 3   long reads: ~100 kb bases / ~1 M samples per read, wide band.  Step = refine_alignment(transitions)
               + estimate_log_likelihoods(wobbling); waves when the DP matrices exceed HBM.
  --config 4   banded-DTW throughput sweep: band width 50 .. 1000 x batch size 1 .. 4096.

Every config carries `cpu_baseline` (the unmodified reference on the host cores, on a stated bounded sample) and a
`parity` block: the GPU's results for the sample's reads are compared with what the reference just computed.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _lists(items):
    return ([it['signal'] for it in items], [it['reference'] for it in items], [it['cb'] for it in items],
            [it['ca'] for it in items], [it['apx'].alignment for it in items])


def _prepare(km, reads, genome, bandwidth):
    """estimator.py:60-74 for every read: slices, contexts, anchors."""
    from nadavca_b200 import synthetic
    from nadavca_b200.genome import Genome
    aligner = synthetic.SyntheticAligner(genome)
    k, cp = km.get_k(), km.get_central_position()
    items = []
    for r in reads:
        apx = aligner.get_signal_alignment(r, bandwidth)
        s0, s1 = apx.signal_range
        a, b = apx.read_sequence_range
        items.append(dict(read=r, apx=apx, signal=r.normalized_signal[s0:s1],
                          reference=Genome.to_numerical(apx.reference_part),
                          cb=Genome.to_numerical(r.sequence[a - cp:a]),
                          ca=Genome.to_numerical(r.sequence[b:b + k - cp - 1])))
    return items


def _timed(fn, steps, warmup, stream, barrier):
    import torch
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1) / steps


def _cpu_jobs(pool, spec, items, bandwidth, mel, kinds):
    """Run the reference on `items` in the worker pool; returns (wall seconds, results per kind)."""
    from oracle import parity
    jobs = []
    for kind, flag in kinds:
        for it in items:
            jobs.append((kind, spec, (it['signal'], it['reference'], it['cb'], it['ca'], it['apx'].alignment,
                                      bandwidth, mel, flag)))
    t0 = time.perf_counter()
    out = pool.map(parity.job, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    res = {}
    for j, (kind, flag) in enumerate(kinds):
        res[(kind, flag)] = out[j * len(items):(j + 1) * len(items)]
    return wall, res


def _parity(km, items, bandwidth, mel, gpu_events, gpu_ll, cpu):
    """Compare the GPU's events / log-likelihoods of the sample reads with the reference's."""
    from oracle import oracle as orc
    from oracle import parity
    om = orc.OracleModel(km.get_k(), km.get_central_position(), 4, km.mean, km.sigma, 'port')
    par = {'reads': len(items), 'alignments_compared': 0, 'event_mismatches': 0, 'tie_accepts': 0, 'max_ll_rel': 0.0,
           'll_mismatches': 0, 'll_rtol': 1e-9}
    for (kind, flag), results in cpu.items():
        for i, (it, want) in enumerate(zip(items, results)):
            if kind == 'refine':
                par['alignments_compared'] += 1
                try:
                    got = parity.compare_events(gpu_events[flag][i], it['signal'], it['reference'], it['cb'], it['ca'],
                                                it['apx'].alignment, bandwidth, mel, om, flag, want=want)
                    par['tie_accepts'] += got == 'tie'
                except AssertionError:
                    par['event_mismatches'] += 1
            else:
                try:
                    rel = parity.ll_max_rel(gpu_ll[i], want)
                    par['max_ll_rel'] = max(par['max_ll_rel'], rel)
                    par['ll_mismatches'] += rel > par['ll_rtol']
                except AssertionError:
                    par['ll_mismatches'] += 1
    return par


def _pool(cores):
    import multiprocessing as mp
    return mp.get_context('spawn').Pool(cores)


def _spec(km):
    from oracle import oracle as orc
    kind = 'reference' if orc.ref_module() is not None else 'port'
    return (km.get_k(), km.get_central_position(), km.mean, km.sigma, 'ref' if kind == 'reference' else 'port'), kind


def _setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    return rank, world, local_rank, torch.device('cuda', local_rank), torch.cuda.current_stream(), barrier


def _finish(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return json.load(fh)['hbm_gbs'], 'measured'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'fallback'


def _base_line(bench, args, world, value, ms, cells_per_s, workload, extra):
    line = {'metric': 'signal_samples_aligned_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'dp_cells_per_sec': cells_per_s,
            'config': workload}
    line.update(extra)
    return line


# ---- configs[0] ------------------------------------------------------------------------------------------------

def run_config0(bench, args):
    import torch
    import nadavca_b200
    from nadavca_b200 import dtw, synthetic
    from nadavca_b200.read import Read
    rank, world, local_rank, dev, stream, barrier = _setup()
    km = bench.load_model()
    km._device = local_rank
    n_reads = args.reads
    genome = synthetic.make_genome(50_000, seed=0)
    make = lambda: [synthetic.make_read(genome, km, 1000 + rank * n_reads + i) for i in range(n_reads)]
    reads = make()
    for r in reads:
        Read.normalize_reads([r])  # per read, align_signal.py:54
    bw, mel = args.bandwidth, bench.DEFAULT_CONFIG['min_event_length']
    items = _prepare(km, reads, genome, bw)
    batch = dtw.Batch(km, *_lists(items), bw, mel)
    samples = int(batch.pack.total_signal)
    cells = batch.cell_counts(True)['refine_transitions']
    width_cells = sum(int((be - bs + 1).sum()) for bs, be in batch.bands())

    def step():
        batch.refine(True, stream)

    with bench.ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            step()
        barrier()
        batch.enable_timing(True)
        clocks.mark()
        ms = _timed(step, args.steps, 0, stream, barrier)
    t = batch.timing()
    batch.enable_timing(False)
    gpu_events = {True: batch.events()[0]}
    # end to end through the C ABI: H2D of every input, refine, D2H of the events and the per-event means
    pk = batch.pack

    def e2e_step():
        b = dtw.Batch.from_pack(km, pk)
        b.refine(True, stream)
        b.events()
        b.event_means()
        b.close()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / 3
    # the public API: align_signal() over fresh copies of the reads (two alignment rounds, host renormalisation)
    cfg = dict(bench.DEFAULT_CONFIG, bandwidth=bw)
    aligner = synthetic.SyntheticAligner(genome)
    list(nadavca_b200.align_signal(None, make(), config=cfg, kmer_model=km, aligner=aligner, reference=genome))
    fresh = make()
    barrier()
    t0 = time.perf_counter()
    out = list(nadavca_b200.align_signal(None, fresh, config=cfg, kmer_model=km, aligner=aligner, reference=genome))
    barrier()
    api_s = time.perf_counter() - t0
    raw_samples = sum(len(r.raw_signal) for r in fresh)
    ms_max = bench.reduce_over_ranks(ms, 'max', dev)
    samples_all = bench.reduce_over_ranks(float(samples), 'sum', dev)
    cells_all = bench.reduce_over_ranks(float(cells), 'sum', dev)
    e2e_max = bench.reduce_over_ranks(e2e_s, 'max', dev)
    api_max = bench.reduce_over_ranks(api_s, 'max', dev)
    raw_all = bench.reduce_over_ranks(float(raw_samples), 'sum', dev)
    if rank == 0:
        hbm, src = _hbm_peak()
        rows_ms = t['rows'][0] / args.steps
        alg_bytes = 2 * (2 * width_cells) * 12  # A and B rows of both directions, 12 B per stored cell
        workload = {'workload': 'configs[0]: nadavca align, %d synthetic reads (~2000 bases, ~20k samples) against a '
                                '50 kb synthetic reference, default config (bandwidth %d, min_event_length 2, '
                                'transitions), kmer_model.hdf5 6-mer' % (n_reads, bw),
                    'reads_per_gpu': n_reads, 'bandwidth': bw,
                    'step': 'refine_alignment(transitions) for the whole batch (one alignment round of align_signal)',
                    'l2': 'DP matrices of %.1f GB per step, larger than L2' % (alg_bytes / 1e9)}
        line = _base_line(bench, args, world, samples_all / (ms_max * 1e-3), ms_max, cells_all / (ms_max * 1e-3),
                          workload, {
            'e2e': {'value': samples_all / e2e_max, 'unit': 'samples/s',
                    'h2d_bytes_per_step': int(sum(getattr(pk, n).nbytes for n in (
                        'signal', 'signal_off', 'reference', 'reference_off', 'context_before', 'context_before_off',
                        'context_after', 'context_after_off', 'anchors', 'anchor_off'))),
                    'd2h_bytes_per_step': int(pk.total_reference * (8 + 8) + pk.n_reads * 4)},
            'api': {'call': 'nadavca_b200.align_signal (two alignment rounds + linear renormalisation)',
                    'value': raw_all / api_max, 'unit': 'raw samples/s', 'seconds': api_max,
                    'aligned': sum(1 for _, res in out if res is not None)},
            'gpu_launches': int(batch.launch_count),
            'clocks': clocks.summary(),
            'stage_ms_per_step': {'rows': rows_ms, 'path': t['path'][0] / args.steps},
            'roofline': {'bound': 'hbm', 'kernel': 'sweep (rows4/rows5)', 'achieved': alg_bytes / (rows_ms * 1e-3) / 1e9,
                         'peak': hbm, 'unit': 'GB/s', 'frac': alg_bytes / (rows_ms * 1e-3) / 1e9 / hbm, 'traffic': None,
                         'peak_source': src,
                         'note': 'latency bound at this batch size: 2 x %d warps for 148 SMs' % n_reads},
        })
        if not args.no_cpu_baseline:
            spec, kind = _spec(km)
            cores = min(os.cpu_count() or 1, n_reads)
            pool = _pool(cores)
            pool.map(len, [[0]] * cores)  # start the workers before timing
            wall, cpu = _cpu_jobs(pool, spec, items, bw, mel, [('refine', True)])
            pool.terminate()
            line['cpu_baseline'] = {'value': samples / wall, 'unit': 'samples/s', 'cores': cores, 'kind': kind,
                                    'sample': 'all %d reads, refine_alignment(transitions), one read per worker, '
                                              '%.1f s wall' % (n_reads, wall)}
            line['parity'] = _parity(km, items, bw, mel, gpu_events, None, cpu)
        print(json.dumps(line))
    batch.close()
    _finish(world)


# ---- configs[3] ------------------------------------------------------------------------------------------------

def run_config3(bench, args):
    import torch
    from nadavca_b200 import dtw, synthetic
    from nadavca_b200.read import Read
    rank, world, local_rank, dev, stream, barrier = _setup()
    km = bench.load_model()
    km._device = local_rank
    n_reads, bases, bw = args.reads, args.bases, args.bandwidth  # defaults: 62 reads (500 over 8 GPUs), 100 kb, 400
    mel = bench.DEFAULT_CONFIG['min_event_length']
    genome = synthetic.make_genome(max(4 * bases, 1_000_000), seed=0)
    reads = [synthetic.make_read(genome, km, 7000 + rank * n_reads + i, n_bases=bases, bandwidth=bw)
             for i in range(n_reads)]
    Read.normalize_reads(reads)
    items = _prepare(km, reads, genome, bw)
    batch = dtw.Batch(km, *_lists(items), bw, mel)
    samples = int(batch.pack.total_signal)
    counts = batch.cell_counts(True)
    cells = counts['refine_transitions'] + counts['estimate_fb'] + counts['estimate_snp']
    width_cells = sum(int((be - bs + 1).sum()) for bs, be in batch.bands())
    free_before = torch.cuda.mem_get_info()[0]

    def step():
        batch.refine(True, stream)
        batch.estimate(True, stream)

    with bench.ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            step()
        barrier()
        batch.enable_timing(True)
        clocks.mark()
        ms = _timed(step, args.steps, 0, stream, barrier)
    t = batch.timing()
    batch.enable_timing(False)
    stage = {name: t[name][0] / args.steps for name in t}
    waves = {name: t[name][1] / args.steps for name in ('rows', 'snp')}
    # separate timings of the two calls (device resident)
    refine_ms = _timed(lambda: batch.refine(True, stream), 1, 0, stream, barrier)
    gpu_events = {True: batch.events()[0]}
    estimate_ms = _timed(lambda: batch.estimate(True, stream), 1, 0, stream, barrier)
    pk = batch.pack

    def e2e_step():
        b = dtw.Batch.from_pack(km, pk)
        b.refine(True, stream)
        b.events()
        b.estimate(True, stream)
        b.log_likelihoods()
        b.close()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    ms_max = bench.reduce_over_ranks(ms, 'max', dev)
    samples_all = bench.reduce_over_ranks(float(samples), 'sum', dev)
    cells_all = bench.reduce_over_ranks(float(cells), 'sum', dev)
    e2e_max = bench.reduce_over_ranks(e2e_s, 'max', dev)
    if rank == 0:
        hbm, src = _hbm_peak()
        mat_bytes = 2 * (2 * width_cells) * 12
        workload = {'workload': 'configs[3]: long reads, %d synthetic reads of ~%d bases / ~%.1f M samples per GPU, '
                                'bandwidth %d (band rows of %d columns), min_event_length 2, kmer_model.hdf5 6-mer'
                                % (n_reads, bases, samples / n_reads / 1e6, bw, 2 * bw + 1),
                    'reads_per_gpu': n_reads, 'bandwidth': bw, 'bases_per_read': bases,
                    'step': 'refine_alignment(transitions) + estimate_log_likelihoods(wobbling)',
                    'l2': 'DP matrices of %.0f GB per step (%.1f GB per read), larger than L2 and HBM: %.0f waves'
                          % (mat_bytes / 1e9, mat_bytes / n_reads / 1e9, waves['rows'] / 2)}
        snp_ms = stage['snp']
        line = _base_line(bench, args, world, samples_all / (ms_max * 1e-3), ms_max,
                          cells_all / (ms_max * 1e-3), workload, {
            'e2e': {'value': samples_all / e2e_max, 'unit': 'samples/s',
                    'h2d_bytes_per_step': int(pk.signal.nbytes + pk.reference.nbytes + pk.anchors.nbytes),
                    'd2h_bytes_per_step': int(pk.total_reference * (8 + 32))},
            'gpu_launches': int(batch.launch_count),
            'clocks': clocks.summary(),
            'stage_ms_per_step': stage,
            'call_ms': {'refine_transitions': refine_ms, 'estimate_wobbling': estimate_ms},
            'cells_per_sec_by_call': {'refine_transitions': counts['refine_transitions'] / (refine_ms * 1e-3),
                                      'estimate_wobbling': (counts['estimate_fb'] + counts['estimate_snp']) /
                                                           (estimate_ms * 1e-3)},
            'memory': {'dp_matrix_bytes_per_read': mat_bytes / n_reads, 'hbm_free_bytes': free_before,
                       'waves_per_call': waves['rows'] / 2},
            'roofline': {'bound': 'hbm', 'kernel': 'snp3_kernel',
                         'achieved': counts['estimate_snp'] * 1.85 / (snp_ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                         'frac': counts['estimate_snp'] * 1.85 / (snp_ms * 1e-3) / 1e9 / hbm, 'traffic': None,
                         'peak_source': src, 'note': 'instruction-issue bound (see DESIGN.md); 1.85 B per SNP cell'},
        })
        if not args.no_cpu_baseline:
            spec, kind = _spec(km)
            cores = min(os.cpu_count() or 1, n_reads, args.cpu_sample or 8)
            pool = _pool(cores)
            pool.map(len, [[0]] * cores)
            wall, cpu = _cpu_jobs(pool, spec, items[:cores], bw, mel, [('refine', True)])
            pool.terminate()
            sample_cells = sum(bench.band_cells_transitions(b) for b in batch.bands()[:cores])
            line['cpu_baseline'] = {'value': sum(len(it['signal']) for it in items[:cores]) / wall, 'unit': 'samples/s',
                                    'cores': cores, 'kind': kind, 'dp_cells_per_sec': sample_cells / wall,
                                    'sample': 'first %d reads, refine_alignment(transitions) only (the estimate of one '
                                              'such read is ~7 core-minutes), one read per worker, %.1f s wall; '
                                              'extrapolate linearly in cells' % (cores, wall)}
            line['parity'] = _parity(km, items[:cores], bw, mel, gpu_events, None, cpu)
        print(json.dumps(line))
    batch.close()
    _finish(world)


# ---- configs[4] ------------------------------------------------------------------------------------------------

def run_config4(bench, args):
    import torch
    from nadavca_b200 import dtw, synthetic
    from nadavca_b200.read import Read
    rank, world, local_rank, dev, stream, barrier = _setup()
    km = bench.load_model()
    km._device = local_rank
    mel = bench.DEFAULT_CONFIG['min_event_length']
    bands = [int(x) for x in args.sweep_bandwidths.split(',')]
    batches = [int(x) for x in args.sweep_batches.split(',')]
    n_max = max(batches)
    genome = synthetic.make_genome(1_000_000, seed=0)
    reads = []
    for i in range(n_max):
        rng = np.random.default_rng(500_000 + rank * n_max + i)
        nb = int(round(rng.normal(args.bases, args.bases / 10.0)))
        reads.append(synthetic.make_read(genome, km, rank * n_max + i, n_bases=nb, bandwidth=max(bands)))
    Read.normalize_reads(reads)
    sweep = []
    check = {}
    with bench.ClockSampler(local_rank) as clocks:
        clocks.mark()
        for bw in bands:
            items = _prepare(km, reads, genome, bw)
            for n in batches:
                sub = items[:n]
                with dtw.Batch(km, *_lists(sub), bw, mel) as batch:
                    samples = int(batch.pack.total_signal)
                    counts = batch.cell_counts(True)
                    steps = max(1, min(args.steps, 3 if n * bw >= 512 * 400 else args.steps))
                    r_ms = _timed(lambda: batch.refine(True, stream), steps, 1, stream, barrier)
                    if n == min(batches, key=lambda x: abs(x - 8)):
                        check[bw] = (sub, batch.events()[0])
                    e_ms = _timed(lambda: batch.estimate(True, stream), steps, 1, stream, barrier)
                    point = {'bandwidth': bw, 'reads': n, 'samples': samples,
                             'refine_transitions_ms': r_ms, 'estimate_wobbling_ms': e_ms,
                             'refine_samples_per_sec': samples / (r_ms * 1e-3),
                             'estimate_samples_per_sec': samples / (e_ms * 1e-3),
                             'refine_cells_per_sec': counts['refine_transitions'] / (r_ms * 1e-3),
                             'estimate_cells_per_sec': (counts['estimate_fb'] + counts['estimate_snp']) / (e_ms * 1e-3)}
                    for key in ('refine_samples_per_sec', 'estimate_samples_per_sec', 'refine_cells_per_sec',
                                'estimate_cells_per_sec'):
                        point[key] = bench.reduce_over_ranks(point[key], 'sum', dev)
                    sweep.append(point)
    if rank == 0:
        hbm, src = _hbm_peak()
        best = max(sweep, key=lambda p: p['estimate_cells_per_sec'])
        ref_point = next((p for p in sweep if p['bandwidth'] == 150 and p['reads'] == max(batches)), sweep[-1])
        workload = {'workload': 'configs[4]: banded-DTW throughput sweep, band width %s x batch size %s reads per GPU '
                                '(~%d bases, ~%dk samples per read), min_event_length 2, kmer_model.hdf5 6-mer'
                                % (bands, batches, args.bases, args.bases // 100),
                    'step': 'refine_alignment(transitions) and estimate_log_likelihoods(wobbling), timed separately, '
                            'inputs resident', 'l2': 'DP matrices larger than L2 from 8 reads up'}
        value = ref_point['refine_samples_per_sec']
        line = _base_line(bench, args, world, value, ref_point['refine_transitions_ms'],
                          ref_point['refine_cells_per_sec'], workload, {
            'value_point': 'refine_alignment(transitions), bandwidth 150, %d reads per GPU' % ref_point['reads'],
            'sweep': sweep, 'clocks': clocks.summary(),
            'roofline': {'bound': 'hbm', 'kernel': 'snp3_kernel (estimate, best point: bandwidth %d, %d reads)'
                                                   % (best['bandwidth'], best['reads']),
                         'achieved': best['estimate_cells_per_sec'] / world * 1.85 / 1e9, 'peak': hbm, 'unit': 'GB/s',
                         'frac': best['estimate_cells_per_sec'] / world * 1.85 / 1e9 / hbm, 'traffic': None,
                         'peak_source': src, 'note': 'instruction-issue bound (see DESIGN.md); 1.85 B per cell'},
        })
        if not args.no_cpu_baseline:
            spec, kind = _spec(km)
            cores = os.cpu_count() or 1
            pool = _pool(cores)
            pool.map(len, [[0]] * cores)
            cpu_rows, par_all = [], {'alignments_compared': 0, 'event_mismatches': 0, 'tie_accepts': 0}
            for bw in bands:
                sub, ev = check[bw]
                sub = sub[:max(1, min(len(sub), cores // len(bands) or 1))]
                wall, cpu = _cpu_jobs(pool, spec, sub, bw, mel, [('refine', True)])
                cells = sum(bench.band_cells_transitions(b) for b in _bands_of(km, sub, bw, mel))
                cpu_rows.append({'bandwidth': bw, 'reads': len(sub), 'wall_s': wall,
                                 'samples_per_sec': sum(len(it['signal']) for it in sub) / wall,
                                 'cells_per_sec': cells / wall})
                par = _parity(km, sub, bw, mel, {True: ev}, None, cpu)
                for key in par_all:
                    par_all[key] += par[key]
            pool.terminate()
            line['cpu_baseline'] = {'value': next(r['samples_per_sec'] for r in cpu_rows if r['bandwidth'] == 150)
                                    if any(r['bandwidth'] == 150 for r in cpu_rows) else cpu_rows[0]['samples_per_sec'],
                                    'unit': 'samples/s', 'cores': cores, 'kind': kind, 'by_bandwidth': cpu_rows,
                                    'sample': 'refine_alignment(transitions) on the first reads of every band width, '
                                              'one read per worker (value: bandwidth 150)'}
            line['parity'] = par_all
        print(json.dumps(line))
    _finish(world)


def _bands_of(km, items, bw, mel):
    from nadavca_b200 import dtw
    with dtw.Batch(km, *_lists(items), bw, mel) as b:
        return b.bands()


def run(bench, args):
    {0: run_config0, 3: run_config3, 4: run_config4}[args.config](bench, args)
