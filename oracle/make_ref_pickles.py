"""Whole-module pickles written by the REAL reference's dnn.DNN, exactly as train_newmethod.py:184,270 /
train.py:179,288 write them (`torch.save(self.dnn, path)`), for the checkpoint-compatibility test: the drop-in
`dnn.DNN` must load them through its __setstate__ (test*.py:37 does `torch.load(path)`).

    python -m oracle.make_ref_pickles          (build container only)

Writes tests/golden/ref_module_{xavier,kaiming}.pth (tiny nets, a few KB each).
"""
import os
import sys

import numpy as np
import torch

from . import jet_oracle as jo

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.path.insert(0, REF)
    import dnn as ref_dnn   # the reference's own module: the pickle records the class as `dnn.DNN`
    sys.path.pop(0)
    for init, act, layers in (("xavier", "tanh", [2, 8, 8, 3]), ("kaiming", "leaky_relu", [3, 6, 5, 4])):
        m = ref_dnn.DNN(layers, 0.0, init)
        flat = jo.make_params(layers, 1234, act, np.float32)
        o = 0
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.from_numpy(flat[o:o + p.numel()]).view_as(p))
                o += p.numel()
        path = os.path.join(OUT, f"ref_module_{init}.pth")
        torch.save(m, path)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
