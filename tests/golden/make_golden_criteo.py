"""Freeze the outputs of the reference's OWN input-side code on a synthetic Criteo TSV.

Run in the build container only (needs /root/reference, which is absent on the GPU box):

    python tests/golden/make_golden_criteo.py

It imports /root/reference/ctr/tfrecord_io.py byte-for-byte under the `tensorflow` shim in oracle/tf_shim (whose
tf.io / tf.train stand-ins keep the records in memory instead of framing them as TFRecords), runs its `build_vocab` and
`write_tfrecord` on two synthetic files (a "train" file the dictionary is built from and a "test" file with unseen
tokens), and stores the TSV bytes, the module's randomly drawn imputation strings, the dictionary and every record the
writer produced in tests/golden/criteo_tsv.npz.  The fixture pins oracle/criteo_oracle.py (CPU test) and the CUDA
parser (GPU test).
"""
import importlib
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle", "tf_shim"))

from oracle import criteo_oracle as C  # noqa: E402


def main():
    sys.path.insert(0, os.path.join(REF, "ctr"))
    ref = importlib.import_module("tfrecord_io")              # draws cat_imputation at import (:11-12)
    import tensorflow as tf_shim
    train = C.synth_tsv(600, seed=4)
    test = C.synth_tsv(97, seed=5, final_newline=False)       # unseen tokens, last line without a newline
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data"))                # the reference pickles to ./data/cat_fea_vocab.pkl (:34)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            for name, text in (("train", train), ("test", test)):
                with open(f"{name}.txt", "wb") as fh:
                    fh.write(text)
            ref.build_vocab("train.txt")
            with open("./data/cat_fea_vocab.pkl", "rb") as fh:
                vocab = pickle.load(fh)
            for name in ("train", "test"):
                ref.write_tfrecord(f"{name}.txt", f"{name}.tfrecord")
                recs = tf_shim.io.WRITTEN[f"{name}.tfrecord"]
                out[f"{name}_int_features"] = np.stack([r["int_features"] for r in recs]).astype(np.float32)
                out[f"{name}_cat_features"] = np.stack([r["cat_features"] for r in recs]).astype(np.int64)
                out[f"{name}_label"] = np.array([r["label"] for r in recs], dtype=np.int64)
                assert recs[0]["int_features"].dtype == np.float32
        finally:
            os.chdir(cwd)
    out["train_tsv"] = np.frombuffer(train, dtype=np.uint8)
    out["test_tsv"] = np.frombuffer(test, dtype=np.uint8)
    out["cat_imputation"] = np.array(ref.cat_imputation)
    out["vocab_tokens"] = np.array(list(vocab.keys()))        # id order (ids are 0..V-1 in insertion order)
    assert list(vocab.values()) == list(range(len(vocab)))
    np.savez_compressed(os.path.join(HERE, "criteo_tsv.npz"), **out)
    print(f"criteo_tsv.npz: {len(vocab)} vocabulary entries, {out['train_label'].size} + {out['test_label'].size} records")


if __name__ == "__main__":
    main()
