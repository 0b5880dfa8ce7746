"""Oracle restatements of the filtered image types (oracle/image_filters.py: Gradient, LoG, Wavelet).  The libraries
behind them (ITK, PyWavelets) are not installable here, so the tests pin the mathematical properties the recalled
algorithms must have; parity with pyradiomics itself stays unpinned."""
import numpy as np
import pytest

from oracle import image_filters as flt


def test_coif1_taps_are_an_orthonormal_filter_bank():
    lo, hi = np.array(flt.COIF1_DEC_LO), np.array(flt.COIF1_DEC_HI)
    assert lo.sum() == pytest.approx(np.sqrt(2), abs=1e-12) and hi.sum() == pytest.approx(0, abs=1e-12)
    assert (lo * lo).sum() == pytest.approx(1, abs=1e-12) and (hi * hi).sum() == pytest.approx(1, abs=1e-12)
    assert (lo * hi).sum() == pytest.approx(0, abs=1e-12)
    for s in (2, 4):  # double-shift orthogonality
        assert (lo[s:] * lo[:-s]).sum() == pytest.approx(0, abs=1e-12)
    np.testing.assert_allclose(hi, [(-1) ** (k + 1) * lo[5 - k] for k in range(6)], atol=1e-15)  # quadrature mirror
    assert (np.arange(6) ** 1 * hi).sum() == pytest.approx(0, abs=1e-10)  # coiflet-1: two vanishing moments
    assert (np.arange(6) ** 0 * hi).sum() == pytest.approx(0, abs=1e-12)


def test_swt_matches_the_published_haar_example():
    # PyWavelets documentation: swt of 1..8 with db1 (dec_lo = [s, s], dec_hi = [-s, s], s = 1/sqrt 2), level 1
    s = 1 / np.sqrt(2)
    x = np.arange(1, 9, dtype=float)
    cA = flt.swt_axis(x[None, :], (s, s), 1)[0]
    cD = flt.swt_axis(x[None, :], (-s, s), 1)[0]
    np.testing.assert_allclose(cA, [2.12132034, 3.53553391, 4.94974747, 6.36396103, 7.77817459, 9.19238816, 10.60660172, 6.36396103], atol=1e-8)
    np.testing.assert_allclose(cD, [-s] * 7 + [4.94974747], atol=1e-8)


def test_swt_level1_energy_and_names():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (12, 10)).astype(np.uint8)
    w2 = flt.wavelet_images(img)
    assert list(w2) == ["wavelet-LH", "wavelet-HL", "wavelet-HH", "wavelet-LL"]
    # undecimated orthonormal bank: sum of the band energies = 2^(#axes) x the image energy (even sizes: no padding)
    e = sum(float((v ** 2).sum()) for v in w2.values())
    assert e == pytest.approx(4 * float((img.astype(float) ** 2).sum()), rel=1e-12)
    w1 = flt.wavelet_images(img, force2D=True, force2Ddimension=0)
    assert list(w1) == ["wavelet-H", "wavelet-L"]
    # first letter <-> x: the 1-D transform of the literal force2D call is the x-only transform
    np.testing.assert_allclose(w1["wavelet-H"], flt.swt_axis(img, flt.COIF1_DEC_HI, 1))
    # a constant image has no detail and a low band of sqrt(2) per axis
    c = flt.wavelet_images(np.full((7, 9), 100, np.uint8))
    assert np.abs(c["wavelet-HH"]).max() < 1e-10 and c["wavelet-LL"] == pytest.approx(200.0)
    assert c["wavelet-LL"].shape == (7, 9)  # odd sizes: padded by a wrapped sample, cropped back


def test_gradient_magnitude():
    yy, xx = np.mgrid[:9, :11]
    g = flt.gradient_image((3 * xx + 4 * yy).astype(np.uint8))
    assert g[1:-1, 1:-1] == pytest.approx(5.0)            # |(3, 4)| in the interior
    assert g[0, 0] == pytest.approx(2.5)                  # ZeroFluxNeumann: one-sided half differences at the border
    assert g.dtype == np.float64 and np.array_equal(g, g.astype(np.float32).astype(np.float64))  # float32 pixels


@pytest.mark.parametrize("sigma", [1.0, 2.0, 3.0])
def test_recursive_gaussian_properties(sigma):
    c0 = flt._deriche_coefficients(sigma, 0)
    c2 = flt._deriche_coefficients(sigma, 2)
    n = 80
    const = np.full((1, n), 7.0)
    np.testing.assert_allclose(flt.recursive_gaussian_lines(const, c0), 7.0, rtol=1e-12)      # unit DC gain, borders included
    np.testing.assert_allclose(flt.recursive_gaussian_lines(const, c2), 0.0, atol=1e-10)      # d2/dx2 of a constant
    x = np.arange(n, dtype=float)
    ramp = flt.recursive_gaussian_lines(x[None, :], c2)[0]
    assert np.abs(ramp[25:55]).max() < 1e-4                                                   # ... and of a ramp (border tails)
    par = flt.recursive_gaussian_lines((x * x)[None, :], c2)[0]
    assert par[40] == pytest.approx(2.0 * sigma * sigma, rel=2e-3)     # d2/dx2 x^2 = 2, scale-normalised by sigma^2
    imp = np.zeros((1, n))
    imp[0, 40] = 1.0
    g = flt.recursive_gaussian_lines(imp, c0)[0]
    want = np.exp(-0.5 * ((x - 40) / sigma) ** 2) / (sigma * np.sqrt(2 * np.pi))
    assert np.abs(g - want).max() < 4e-3 * want.max() + 2e-4          # Deriche's 4th-order fit of the Gaussian
    assert g.sum() == pytest.approx(1.0, rel=1e-7)
    assert np.abs(g[40 - 10:40] - g[40 + 10:40:-1]).max() < 1e-12      # symmetric response


def test_log_image_of_a_blob_and_names():
    yy, xx = np.mgrid[:48, :40]
    blob = (200 * np.exp(-((yy - 24.0) ** 2 + (xx - 20.0) ** 2) / (2 * 3.0 ** 2))).astype(np.uint8)
    out = flt.log_image(blob, 3.0)
    assert out.shape == blob.shape and np.array_equal(out, out.astype(np.float32).astype(np.float64))
    assert out[24, 20] == out.min() and out[24, 20] < -50     # bright blob: strongly negative Laplacian at its centre
    assert abs(out[2, 2]) < 1.0
    assert flt.log_name(1.0) == "log-sigma-1-0-mm-3D" and flt.log_name(2.5) == "log-sigma-2-5-mm-3D"
    assert list(flt.filtered_images(blob, "LoG", {"sigma": [1.0, 2.0]})) == ["log-sigma-1-0-mm-3D", "log-sigma-2-0-mm-3D"]
    with pytest.raises(ValueError):
        flt.log_image(np.zeros((3, 10), np.uint8), 1.0)
