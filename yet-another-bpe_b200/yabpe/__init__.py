"""yabpe -- B200-native byte-level BPE (train_bpe / Tokenizer), drop-in for the hot path of
DreamOneX/yet-another-bpe.  Public names mirror /root/reference/src/yet_another_bpe/__init__.py."""
from .tokenizer import BBPETokenizer, Tokenizer
from .trainer import BBPEModel, BBPETrainer, BBPETrainerConfig, train_bpe

__all__ = ["BBPETokenizer", "BBPETrainer", "BBPETrainerConfig", "BBPEModel", "Tokenizer", "train_bpe"]
