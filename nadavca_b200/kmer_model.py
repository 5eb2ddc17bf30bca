"""k-mer model loading (reference nadavca/kmer_model.py:5-29) without h5py."""
import numpy as np

from . import hdf5_mini
from .alphabet import alphabet, inv_alphabet
from .dtw import KmerModel


def kmer_to_id(kmer):
    """Base-4 number of the k-mer, first base most significant (kmer_model.py:5-10)."""
    result = 0
    for base in kmer:
        result = result * len(alphabet) + inv_alphabet[base]
    return result


def load_kmer_model(filename):
    """Root attribute ``central_pos`` and dataset ``model`` with rows (kmer, mean, sd) (kmer_model.py:13-27)."""
    file = hdf5_mini.File(filename)
    central_position = int(file.attrs['central_pos'])
    table = file['model']
    names = table.dtype.names
    kmers = table[names[0]]
    k = len(kmers[0])
    mean = np.zeros(len(table))
    sigma = np.zeros(len(table))
    codes = np.frombuffer(kmers.tobytes(), dtype=np.uint8).reshape(len(table), k)
    lut = np.full(256, -1, dtype=np.int64)
    for base, index in inv_alphabet.items():
        lut[ord(base)] = index
    digits = lut[codes]
    if digits.min() < 0:
        raise KeyError('k-mer model contains a base outside ACGT')
    ids = np.zeros(len(table), dtype=np.int64)
    for col in range(k):
        ids = ids * len(alphabet) + digits[:, col]
    mean[ids] = table[names[1]]
    sigma[ids] = table[names[2]]
    return KmerModel(k, central_position, len(alphabet), mean, sigma)


KmerModel.load_from_hdf5 = staticmethod(load_kmer_model)
