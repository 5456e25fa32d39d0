"""oracle/make_golden_encoder.py -- TEST INFRASTRUCTURE ONLY.

Writes tests/golden/encoder_small.npz: token ids + the fp32 embeddings the real
transformers MPNetModel (random init, seed 0, 12 layers) with the restated
sentence-transformers pooling produces for them, in two weight settings (HF default
init; biases / LayerNorm / relative bias randomised).  The weights themselves are
regenerated from the seed by the tests (same torch build in the same image).

    python -m oracle.make_golden_encoder
"""
from pathlib import Path

import numpy as np

from oracle import encoder_oracle as eo

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
LENGTHS = [2, 3, 9, 31, 64, 65, 127, 200, 384]


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    seqs = eo.synthetic_ids(len(LENGTHS), LENGTHS, seed=7)
    ids = np.concatenate([np.asarray(s, np.int32) for s in seqs])
    cu = np.zeros(len(seqs) + 1, np.int32)
    cu[1:] = np.cumsum([len(s) for s in seqs])
    plain = eo.st_encode_ids(eo.build_model(seed=0, perturb=False), seqs)
    pert = eo.st_encode_ids(eo.build_model(seed=0, perturb=True), seqs)
    pert_raw = eo.st_encode_ids(eo.build_model(seed=0, perturb=True), seqs, normalize=False)
    np.savez_compressed(OUT / "encoder_small.npz", ids=ids, cu_seqlens=cu, emb_plain=plain, emb_perturbed=pert,
                        emb_perturbed_unnormalized=pert_raw)
    print("wrote encoder_small.npz", plain.shape, float(np.abs(plain).max()))


if __name__ == "__main__":
    main()
