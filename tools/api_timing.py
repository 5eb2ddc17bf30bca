"""Where the time of the PUBLIC API goes (estimate_snps / align_signal on synthetic reads): cProfile summary."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nadavca_b200
from nadavca_b200 import synthetic
from nadavca_b200.kmer_model import KmerModel
def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    km = KmerModel.load_from_hdf5(os.path.join(ROOT, 'nadavca_b200', 'default', 'kmer_model.hdf5'))
    cfg = dict(bandwidth=150, snp_prior_probability=0.001, min_event_length=2, model_wobbling=True, model_transitions=True,
               tweak_signal_normalization=True, normalization_event_length=10)
    genome = synthetic.make_genome(4_600_000, seed=0)
    aligner = synthetic.SyntheticAligner(genome)
    reads = [synthetic.make_read(genome, km, i) for i in range(R)]
    samples = sum(len(r.raw_signal) for r in reads)
    nadavca_b200.estimate_snps(None, reads, reference=genome, config=cfg, kmer_model=km, independent=True, aligner=aligner)  # warm-up: worker pool, workspaces
    for name, fn in (('estimate_snps(independent=True)', lambda: nadavca_b200.estimate_snps(None, reads, reference=genome, config=cfg, kmer_model=km, independent=True, aligner=aligner)),
                     ('align_signal', lambda: list(nadavca_b200.align_signal(None, reads, config=cfg, kmer_model=km, aligner=aligner, reference=genome)))):
        pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable(); fn(); pr.disable(); dt = time.perf_counter() - t0
        print('== %s: %d reads, %.2f s, %.2f M raw samples/s' % (name, R, dt, samples / dt / 1e6))
        st = pstats.Stats(pr); st.sort_stats('cumulative')
        import io; buf = io.StringIO(); pstats.Stats(pr, stream=buf).sort_stats('tottime').print_stats(16); print('\n'.join(buf.getvalue().splitlines()[6:26]))
        buf = io.StringIO(); pstats.Stats(pr, stream=buf).sort_stats('cumulative').print_stats(22); print('\n'.join(buf.getvalue().splitlines()[6:32]))


if __name__ == '__main__':
    main()
