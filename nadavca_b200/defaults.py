"""Default paths and constants (reference nadavca/defaults.py:4-8).

The reference's default model path names ``10kmer_fact2.h5``, which is not shipped with it; the shipped 6-mer
``kmer_model.hdf5`` (the file BASELINE.json pins) is the default here.
"""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
KMER_MODEL_FILE = os.path.join(_HERE, 'default', 'kmer_model.hdf5')
CONFIG_FILE = os.path.join(_HERE, 'default', 'config.yaml')
BWA_EXECUTABLE = 'bwa'
GROUP_NAME = 'Analyses/Basecall_1D_000'
RENORM_ROUNDS = 3

CONFIG_KEYS = ('bandwidth', 'snp_prior_probability', 'min_event_length', 'model_wobbling', 'model_transitions',
               'tweak_signal_normalization', 'normalization_event_length')


def load_config(config=CONFIG_FILE):
    """Config is a dict or a YAML path (reference estimate_snps.py:21-27; ``yaml.safe_load`` instead of the bare
    ``yaml.load`` that PyYAML >= 6 rejects)."""
    if isinstance(config, dict):
        cfg = dict(config)
    else:
        import yaml
        with open(config, 'r') as file:
            cfg = yaml.safe_load(file)
    missing = [key for key in CONFIG_KEYS if key not in cfg]
    if missing:
        raise KeyError('config is missing keys: {}'.format(', '.join(missing)))
    return cfg
