#!/usr/bin/env python
"""Settles U1 (oracle/U1_ANGLES.md) on a machine where pyradiomics IS installed: runs the reference's literal
call pattern (2-D SimpleITK image, force2D: True) on the vertical-stripes fixture and reports which angle
reading the installed pyradiomics follows.  Not used by tests (pyradiomics is absent from this image)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    with open(os.path.join(ROOT, "tests", "golden", "u1_stripes.json")) as fh:
        fx = json.load(fh)
    try:
        import SimpleITK as sitk
        from radiomics import featureextractor
    except ImportError as e:
        print("pyradiomics / SimpleITK not importable (%s): U1 stays unsettled" % e)
        return 2
    img = sitk.GetImageFromArray(np.array(fx["image"], dtype=np.uint8))   # RadiomicExtractor.py:31
    msk = sitk.GetImageFromArray(np.array(fx["mask"], dtype=np.uint8))    # RadiomicExtractor.py:36
    ex = featureextractor.RadiomicsFeatureExtractor(label=255, binWidth=10, force2D=True, additionalInfo=False)
    ex.disableAllFeatures()
    ex.enableFeatureClassByName("glcm")
    got = float(ex.execute(img, msk, label=255)["original_glcm_JointEnergy"])
    lit, inp = fx["JointEnergy_literal"], fx["JointEnergy_inplane"]
    print("pyradiomics JointEnergy = %.12g; literal reading %.12g, in-plane reading %.12g" % (got, lit, inp))
    print("=> pyradiomics follows the %s reading" % ("LITERAL (1 angle)" if abs(got - lit) < abs(got - inp) else "IN-PLANE (4 angles)"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
