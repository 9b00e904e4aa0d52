"""First 128 pairs of the reference's shipped human random-split test set as a small fixture.

    python tests/golden/make_human_fixture.py        # needs /root/reference (build container only)

BASELINE.json's configs[0] / north_star ask for AUROC / AUPRC parity on this split.  RDKit, DGL and
the ESM / ChemBERTa checkpoints are absent here, so tests/test_auroc_parity_gpu.py featurises the
pairs with the surrogate of druglamp_b200.synth.batch_from_records (real residue tokens and real
labels; atom count from the SMILES string; seeded synthetic graph topology and embeddings) and
feeds the SAME tensors to the oracle and to the product.
"""
import csv
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/datasets/human/random/test.csv"

if __name__ == "__main__":
    rows = []
    with open(SRC) as f:
        for i, r in enumerate(csv.DictReader(f)):
            if i >= 128:
                break
            rows.append({"smiles": r["SMILES"], "protein": r["Protein"], "y": int(float(r["Y"]))})
    with open(os.path.join(HERE, "human_random_test_128.json"), "w") as f:
        json.dump(rows, f)
    print(len(rows), "pairs,", sum(r["y"] for r in rows), "positives")
