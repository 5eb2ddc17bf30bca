"""Nucleotide alphabet tables (reference nadavca/alphabet.py:1-4)."""
alphabet = ['A', 'C', 'G', 'T']
inv_alphabet = {base: index for index, base in enumerate(alphabet)}
complement = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
numerical_complement = {index: inv_alphabet[complement[base]] for index, base in enumerate(alphabet)}
