"""estimate_snps() -- reference nadavca/estimate_snps.py:13-70, batched on the GPU."""
import sys

from . import defaults
from .alignment import ApproximateAligner
from .estimator import ProbabilityEstimator
from .genome import Genome
from .kmer_model import KmerModel
from .read import Read


def estimate_snps(reference_filename,
                  reads,
                  reference=None,
                  config=defaults.CONFIG_FILE,
                  kmer_model=defaults.KMER_MODEL_FILE,
                  bwa_executable=defaults.BWA_EXECUTABLE,
                  independent=False,
                  group_name=defaults.GROUP_NAME,
                  aligner=None,
                  process_group=None):
    """Same arguments as the reference plus `aligner` (any object with get_signal_alignment(read, bandwidth);
    BWA mapping itself is out of scope) and `process_group` (consensus mode over several GPUs: every rank passes
    its own shard of reads).  Returns a list: with `independent` one entry per input read in input order (its
    ``Chunk``, or None for a read without an alignment), otherwise one ``Chunk`` per overlap group sorted by start."""
    if aligner is None:
        raise ValueError('estimate_snps needs an `aligner` (any object with get_signal_alignment(read, bandwidth)): '
                         'mapping reads with BWA is host work outside nadavca_b200')
    try:
        config = defaults.load_config(config)
    except FileNotFoundError:
        sys.stderr.write('failed to load config: {} not found\n'.format(config))
        return None
    if isinstance(kmer_model, str):
        try:
            kmer_model = KmerModel.load_from_hdf5(kmer_model)
        except FileNotFoundError:
            sys.stderr.write('failed to load k-mer model: {} not found\n'.format(kmer_model))
            return None
    if reference is None:
        try:
            reference = Genome.load_from_fasta(reference_filename)[0].bases
        except FileNotFoundError:
            sys.stderr.write("failed to process: reference {} doesn't exist\n".format(reference_filename))
            return None
    if aligner is None:
        aligner = ApproximateAligner(bwa_executable, reference, reference_filename)
    estimator = ProbabilityEstimator(kmer_model, aligner, config)

    reads = list(reads)
    for i, read in enumerate(reads):
        if isinstance(read, str):
            reads[i] = Read.load_from_fast5(read, group_name)
    # ONE median/MAD pooled over all reads (estimate_snps.py:61) -- over the reads of all ranks in a sharded job
    # ... on the device: exact radix select of the two medians + clip kernel (csrc/select.cu)
    Read.normalize_reads(reads, process_group, device=kmer_model.device)
    return estimator.estimate_probabilities(reference, reads, independent=independent,
                                            process_group=process_group)
