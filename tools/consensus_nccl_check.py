"""Consensus mode over several GPUs: every rank runs estimate_snps(independent=False) on ITS shard of the reads with
the NCCL process group; the result must equal what one GPU computes from all reads (rank 0 checks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/consensus_nccl_check.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import nadavca_b200
from nadavca_b200 import synthetic
from nadavca_b200.estimator import shard_reads
from nadavca_b200.kmer_model import KmerModel

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
km = KmerModel.load_from_hdf5(os.path.join(ROOT, 'nadavca_b200', 'default', 'kmer_model.hdf5'))
cfg = dict(bandwidth=150, snp_prior_probability=0.001, min_event_length=2, model_wobbling=True, model_transitions=True,
           tweak_signal_normalization=True, normalization_event_length=10)
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 64
G = int(sys.argv[2]) if len(sys.argv) > 2 else 30_000
n_bases = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
genome = synthetic.make_genome(G, seed=5)
make = lambda: [synthetic.make_read(genome, km, 2000 + i, n_bases=n_bases) for i in range(n_reads)]
aligner = synthetic.SyntheticAligner(genome)
reads = make()
work = [len(r.raw_signal) for r in reads]
mine = shard_reads(work, world)[rank]
# the public API on this rank's shard: pooled normalisation over the reads of ALL ranks (exact distributed median),
# refine -> tweak -> SNP DP on the local shard, one NCCL all-reduce of the per-position sums, posterior
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
chunks = nadavca_b200.estimate_snps(None, [reads[i] for i in mine], reference=genome, config=cfg, kmer_model=km,
                                    independent=False, aligner=aligner, process_group=dist.group.WORLD)
torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
if rank == 0:
    single = nadavca_b200.estimate_snps(None, make(), reference=genome, config=cfg, kmer_model=km, independent=False,
                                        aligner=aligner)
    assert [(c.start, c.end) for c in chunks] == [(c.start, c.end) for c in single]
    worst = 0.0
    for a, b in zip(chunks, single):
        assert np.array_equal(a.coverage, b.coverage)
        np.testing.assert_allclose(a.values, b.values, rtol=1e-9, atol=1e-14)
        worst = max(worst, float(np.abs(a.values - b.values).max()))
    print('consensus over %d GPUs: %d reads, %d groups, %d positions, max |diff| vs one GPU %.2e, %.3f s' %
          (world, n_reads, len(chunks), sum(c.end - c.start for c in chunks), worst, t1 - t0))
dist.barrier()
dist.destroy_process_group()
