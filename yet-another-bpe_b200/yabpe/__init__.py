"""yabpe -- B200-native byte-level BPE (train_bpe / Tokenizer), drop-in for the hot path of
DreamOneX/yet-another-bpe.  Public names mirror /root/reference/src/yet_another_bpe/__init__.py."""
__version__ = "0.1.0"          # the reference package's version attribute (src/yet_another_bpe/__init__.py:3)

from .tokenizer import BBPETokenizer, Tokenizer
from .trainer import BBPEModel, BBPETrainer, BBPETrainerConfig, train_bpe

__all__ = ["__version__", "BBPETokenizer", "BBPETrainer", "BBPETrainerConfig", "BBPEModel", "Tokenizer", "train_bpe"]
