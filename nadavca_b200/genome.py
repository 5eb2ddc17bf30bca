"""FASTA / FASTQ containers and base conversions (reference nadavca/genome.py)."""
import numpy

from .alphabet import complement, inv_alphabet

_LUT = numpy.full(256, -1, dtype=numpy.int32)
for _base, _index in inv_alphabet.items():
    _LUT[ord(_base)] = _index
_COMP = numpy.arange(256, dtype=numpy.uint8)
_HAS_COMP = numpy.zeros(256, dtype=bool)
for _base, _other in complement.items():
    _COMP[ord(_base)] = ord(_other)
    _HAS_COMP[ord(_base)] = True


def _as_bytes(sequence):
    if isinstance(sequence, str):
        return numpy.frombuffer(sequence.encode('ascii'), dtype=numpy.uint8)
    arr = numpy.asarray(sequence)
    if arr.dtype.kind == 'U':
        if arr.size == 0:
            return numpy.zeros(0, dtype=numpy.uint8)
        if arr.dtype.itemsize == 4:  # arrays of single characters: read the UCS4 code points directly
            codes = numpy.ascontiguousarray(arr).view(numpy.uint32)
            if codes.max() < 256:
                return codes.astype(numpy.uint8)
        return arr.astype('S1').view(numpy.uint8)
    if arr.dtype.kind == 'S':
        return arr.view(numpy.uint8)
    return arr.astype(numpy.uint8)


class Genome:
    def __init__(self, desc_line=None):
        self.description = desc_line
        self.lines = []
        self.bases = None

    def _finish(self):
        self.bases = numpy.array(list(''.join(self.lines)))

    @staticmethod
    def to_numerical(sequence):
        """Bases -> int array, A,C,G,T = 0..3 (genome.py:16-17); any other symbol raises KeyError like the reference."""
        codes = _LUT[_as_bytes(sequence)]
        if codes.size and codes.min() < 0:
            bad = numpy.asarray(sequence)[int(numpy.argmin(codes))] if not isinstance(sequence, str) else \
                sequence[int(numpy.argmin(codes))]
            raise KeyError(bad)
        return codes.astype(int)

    @staticmethod
    def reverse_complement(sequence):
        """genome.py:19-23: complement every base and reverse; returns an array of 1-char strings."""
        raw = _as_bytes(sequence)
        known = _HAS_COMP[raw]
        if not known.all():  # the reference's dict lookup raises KeyError on the first base it does not know
            raise KeyError(chr(int(raw[int(numpy.argmin(known))])))
        return _COMP[raw][::-1].astype(numpy.uint32).view('U1')  # UCS4 code points -> array of 1-char strings

    @staticmethod
    def load_from_fasta(filename):
        result = []
        with open(filename, 'r') as file:
            current = None
            for line in file:
                if line.startswith('>'):
                    current = Genome(line.rstrip())
                    result.append(current)
                elif current is not None:
                    current.lines.append(line.rstrip())
        for genome in result:
            genome._finish()
        return result

    @staticmethod
    def create_from_fastq_string(fastq_string):
        result = []
        expect_sequence = False
        for line in fastq_string.split('\n'):
            if len(line) > 0 and line[0] == '@':
                result.append(Genome(line))
                expect_sequence = True
            elif expect_sequence:
                result[-1].lines.append(line.rstrip())
                expect_sequence = False
        for genome in result:
            genome._finish()
        return result

    @staticmethod
    def load_from_fastq(filename):
        with open(filename, 'r') as file:
            return Genome.create_from_fastq_string(file.read())
