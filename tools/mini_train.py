import sys
sys.path.insert(0, '/root/repo/yet-another-bpe_b200')
import torch
import yabpe
for i in range(3):
    v, m = yabpe.train_bpe('/root/repo/tests/fixtures_gpt2/corpus.en', 500, ['<|endoftext|>'])
    torch.cuda.synchronize()
    print("run", i, len(m), flush=True)
