"""Regenerates tests/golden/*.  The reference's engine (pyradiomics 3.1.0) is not installable
here, so the golden vectors are (a) the integer matrices of pyradiomics' public docstring
examples typed in by hand, and (b) ORACLE outputs on seeded synthetic patches
(they pin the oracle against regressions; they are not outputs of the reference itself).
Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from multimodal_isic_b200 import synth
    from oracle import radiomics_oracle as orc

    a10 = {
        "source": "pyradiomics docstring examples (glcm.py, glrlm.py, glszm.py, gldm.py, ngtdm.py); SURVEY.md A.10",
        "I1": [[1, 2, 5, 2, 3], [3, 2, 1, 3, 1], [1, 3, 5, 5, 2], [1, 1, 1, 1, 2], [1, 2, 4, 3, 5]],
        "glcm_sym_angle01": [[6, 4, 3, 0, 0], [4, 0, 2, 1, 3], [3, 2, 0, 1, 2], [0, 1, 1, 0, 0], [0, 3, 2, 0, 2]],
        "I2": [[5, 2, 5, 4, 4], [3, 3, 3, 1, 3], [2, 1, 1, 1, 3], [4, 2, 2, 2, 3], [3, 5, 3, 3, 2]],
        "glrlm_angle01": [[1, 0, 1, 0, 0], [3, 0, 1, 0, 0], [4, 1, 1, 0, 0], [1, 1, 0, 0, 0], [3, 0, 0, 0, 0]],
        "glszm_8conn": [[0, 0, 0, 1, 0], [1, 0, 0, 0, 1], [1, 0, 1, 0, 1], [1, 1, 0, 0, 0], [3, 0, 0, 0, 0]],
        "gldm_alpha0_8nb": [[0, 1, 2, 1], [1, 2, 3, 0], [1, 4, 4, 0], [1, 2, 0, 0], [3, 0, 0, 0]],
        "I3": [[1, 2, 5, 2], [3, 5, 1, 3], [1, 3, 5, 5], [3, 1, 1, 1]],
        "ngtdm_n": [6, 2, 4, 0, 4],
        "ngtdm_s": [13.35, 2.0, 3.0333333333333334, 0.0, 10.075],
    }
    with open(os.path.join(HERE, "a10_matrices.json"), "w") as fh:
        json.dump(a10, fh, indent=1)
    imgs, masks = synth.make_patches(6, 64, seed=0)
    rows = {}
    for name, st in (("inplane_bw10", dict(label=255, binWidth=10, force2D=False)),
                     ("literal_bw10", dict(label=255, binWidth=10, force2D=True)),
                     ("inplane_bw25", dict(label=255, binWidth=25, force2D=False))):
        rows[name] = np.array([list(orc.execute(imgs[b], masks[b], st).values()) for b in range(len(imgs))])
    np.savez_compressed(os.path.join(HERE, "oracle_features_seed0.npz"), images=imgs, masks=masks,
                        names=np.array(orc.feature_names()), **rows)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
