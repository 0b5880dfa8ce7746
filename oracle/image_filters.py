"""Filtered image types of the pyradiomics parameter file -- Gradient, LoG, Wavelet.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  ``/root/reference/params.yml:137-145`` enables ``Wavelet``, ``LoG`` (sigma 1, 2, 3) and
``Gradient`` next to the point-wise types; pyradiomics 3.1.0 (``imageoperations.getWaveletImage`` /
``getLoGImage`` / ``getGradientImage``) delegates them to PyWavelets and SimpleITK/ITK, none of which is
installable here.  Each function below restates the *published algorithm* of the library routine it names, in
float64 NumPy with the library's own operation order (no fused multiply-adds), so that the CUDA kernels can be
compared with it bit for bit.  What cannot be checked here is that the recalled library internals are exact
(filter taps, Deriche coefficients, boundary rules): the tests pin their mathematical properties instead
(orthonormal taps and perfect reconstruction; unit DC gain, zero response to constants, d2/dx2 of a parabola).

Array convention: 2-D ``(y, x)`` arrays as produced by ``sitk.GetArrayFromImage``; ITK's dimension 0 is x.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

# ---------------------------------------------------------------------------------------------- wavelet (PyWavelets)
# pywt.Wavelet('coif1').dec_lo / dec_hi (pyradiomics default ``wavelet: 'coif1'``)
COIF1_DEC_LO = (-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554,
                0.3378976624578092, -0.0727326195128539)
COIF1_DEC_HI = (0.0727326195128539, 0.3378976624578092, -0.8525720202122554, 0.38486484686420286,
                0.0727326195128539, -0.01565572813546454)


def swt_axis(x, taps, axis):
    """One level of the stationary (undecimated) wavelet transform along ``axis`` with periodic extension:
    PyWavelets ``swt_axis`` -> C ``downsampling_convolution_periodization(step=1)``:
    ``out[n] = sum_j taps[j] * x[(n + F/2 - j) mod N]``, accumulated in tap order."""
    x = np.asarray(x, dtype=np.float64)
    F = len(taps)
    out = np.zeros_like(x)
    for j in range(F):
        out = out + taps[j] * np.roll(x, j - F // 2, axis=axis)  # roll by s: out[n] = x[n - s]
    return out


def wavelet_images(image, force2D=False, force2Ddimension=0):
    """pyradiomics ``imageoperations.getWaveletImage`` / ``_swt3`` (level 1, start_level 0, coif1) for a 2-D image.
    Returns ``OrderedDict[name -> float64 array]`` in pyradiomics' yield order.  The transform axes are
    ``range(Nd - 1, -1, -1)`` minus ``force2Ddimension`` when ``force2D`` (the same axis removal as the texture
    angles, oracle/U1_ANGLES.md): the reference's literal call gives a 1-D transform along x -> ``wavelet-H``,
    ``wavelet-L``; without force2D: ``wavelet-LH, -HL, -HH, -LL`` (first letter <-> x)."""
    a = np.asarray(image).astype(np.float64)
    H, W = a.shape
    axes = [1, 0]
    if force2D:
        axes.remove(int(force2Ddimension))
    # odd sizes are padded by one wrapped sample (numpy.pad 'wrap') and cropped afterwards
    data = np.pad(a, [(0, 1 if H % 2 else 0), (0, 1 if W % 2 else 0)], mode="wrap")
    dec = {"": data}
    for ax in axes:  # pywt.swtn applies the filters axis by axis; key letter k <-> axes[k]
        nxt = OrderedDict()
        for key, arr in dec.items():
            nxt[key + "a"] = swt_axis(arr, COIF1_DEC_LO, ax)
            nxt[key + "d"] = swt_axis(arr, COIF1_DEC_HI, ax)
        dec = nxt
    out = OrderedDict()
    approx_key = "a" * len(axes)
    keys = sorted(dec.keys())  # itertools.product('ad', ...) order: 'aa', 'ad', 'da', 'dd'
    for key in keys:
        if key != approx_key:
            out["wavelet-" + key.replace("a", "L").replace("d", "H")] = dec[key][:H, :W]
    out["wavelet-" + "L" * len(axes)] = dec[approx_key][:H, :W]
    return out


# ---------------------------------------------------------------------------------------------- gradient (ITK)
def gradient_image(image):
    """pyradiomics ``getGradientImage`` -> ``sitk.GradientMagnitudeImageFilter`` (``gradientUseSpacing`` True, spacing
    (1, 1)): central differences ``0.5 * (f[i+1] - f[i-1])`` per axis with the ZeroFluxNeumann boundary condition
    (out-of-range neighbours take the border value), magnitude in double, result cast to float32 (SimpleITK's
    output pixel type).  Returned as float64 holding the float32 values."""
    a = np.asarray(image).astype(np.float64)
    p = np.pad(a, 1, mode="edge")
    dy = 0.5 * (p[2:, 1:-1] - p[:-2, 1:-1])
    dx = 0.5 * (p[1:-1, 2:] - p[1:-1, :-2])
    return np.sqrt(dx * dx + dy * dy).astype(np.float32).astype(np.float64)


# ---------------------------------------------------------------------------------------------- LoG (ITK)
def _deriche_coefficients(sigma, order, normalize_across_scale=True, spacing=1.0):
    """ITK ``RecursiveGaussianImageFilter::SetUp`` (Deriche's 4th-order recursive approximation of the Gaussian
    and its derivatives): returns N0..N3, D1..D4, M1..M4, BN1..BN4, BM1..BM4 as a dict of Python floats."""
    sd = sigma / spacing
    W1, L1, W2, L2 = 0.6681, -1.3932, 2.0787, -1.3732
    A1 = (1.3530, -0.6724, -1.3563)
    B1 = (1.8151, -3.4327, 5.2318)
    A2 = (-0.3531, 0.6724, 0.3446)
    B2 = (0.0902, 0.6100, -2.2355)
    c1, c2 = np.cos(W1 / sd), np.cos(W2 / sd)
    s1, s2 = np.sin(W1 / sd), np.sin(W2 / sd)
    e1, e2 = np.exp(L1 / sd), np.exp(L2 / sd)

    def ncoef(a1, b1, a2, b2):
        n0 = a1 + a2
        n1 = e2 * (b2 * s2 - (a2 + 2 * a1) * c2)
        n1 += e1 * (b1 * s1 - (a1 + 2 * a2) * c1)
        n2 = (a1 + a2) * c2 * c1
        n2 -= b1 * c2 * s1 + b2 * c1 * s2
        n2 *= 2 * e1 * e2
        n2 += a2 * e1 * e1 + a1 * e2 * e2
        n3 = e2 * e1 * e1 * (b2 * s2 - a2 * c2)
        n3 += e1 * e2 * e2 * (b1 * s1 - a1 * c1)
        return n0, n1, n2, n3, n0 + n1 + n2 + n3, n1 + 2 * n2 + 3 * n3, n1 + 4 * n2 + 9 * n3

    D4 = e1 * e1 * e2 * e2
    D3 = -2 * c1 * e1 * e2 * e2
    D3 += -2 * c2 * e2 * e1 * e1
    D2 = 4 * c2 * c1 * e1 * e2
    D2 += e1 * e1 + e2 * e2
    D1 = -2 * (e2 * c2 + e1 * c1)
    SD = 1.0 + D1 + D2 + D3 + D4
    DD = D1 + 2 * D2 + 3 * D3 + 4 * D4
    ED = D1 + 4 * D2 + 9 * D3 + 16 * D4
    if order == 0:
        n0, n1, n2, n3, SN, DN, EN = ncoef(A1[0], B1[0], A2[0], B2[0])
        alpha0 = 2 * SN / SD - n0
        N = [n0 / alpha0, n1 / alpha0, n2 / alpha0, n3 / alpha0]
        symmetric = True
    elif order == 2:
        scale = sigma * sigma if normalize_across_scale else 1.0
        a0, a1_, a2_, a3, SN0, DN0, EN0 = ncoef(A1[0], B1[0], A2[0], B2[0])
        b0, b1_, b2_, b3, SN2, DN2, EN2 = ncoef(A1[2], B1[2], A2[2], B2[2])
        beta = -(2 * SN2 - SD * b0) / (2 * SN0 - SD * a0)
        n0, n1, n2, n3 = b0 + beta * a0, b1_ + beta * a1_, b2_ + beta * a2_, b3 + beta * a3
        SN, DN, EN = SN2 + beta * SN0, DN2 + beta * DN0, EN2 + beta * EN0
        alpha2 = (EN * SD * SD - ED * SN * SD - 2 * DN * DD * SD + 2 * DD * DD * SN) / (SD * SD * SD)
        N = [n0 * (scale / alpha2), n1 * (scale / alpha2), n2 * (scale / alpha2), n3 * (scale / alpha2)]
        symmetric = True
    else:
        raise NotImplementedError("only the zero- and second-order filters are needed for LoG")
    D = [D1, D2, D3, D4]
    sign = 1.0 if symmetric else -1.0
    M = [sign * (N[1] - D1 * N[0]), sign * (N[2] - D2 * N[0]), sign * (N[3] - D3 * N[0]), -sign * D4 * N[0]]
    SNs = N[0] + N[1] + N[2] + N[3]
    SMs = M[0] + M[1] + M[2] + M[3]
    SDs = 1.0 + D1 + D2 + D3 + D4
    BN = [d * SNs / SDs for d in D]
    BM = [d * SMs / SDs for d in D]
    return dict(N=[float(v) for v in N], D=[float(v) for v in D], M=[float(v) for v in M],
                BN=[float(v) for v in BN], BM=[float(v) for v in BM])


def recursive_gaussian_lines(data, coef):
    """ITK ``RecursiveSeparableImageFilter::FilterDataArray`` on every row of ``data`` ([lines, ln] float64, ln >= 4):
    causal + anti-causal 4th-order recursions with ITK's border initialisation, products and sums in ITK's order."""
    x = np.asarray(data, dtype=np.float64)
    ln = x.shape[1]
    if ln < 4:
        raise ValueError("ITK's recursive filters need at least 4 pixels along the filtered direction")
    N0, N1, N2, N3 = coef["N"]
    D1, D2, D3, D4 = coef["D"]
    M1, M2, M3, M4 = coef["M"]
    BN1, BN2, BN3, BN4 = coef["BN"]
    BM1, BM2, BM3, BM4 = coef["BM"]
    s1 = np.zeros_like(x)
    v = x[:, 0]
    s1[:, 0] = v * N0 + v * N1 + v * N2 + v * N3
    s1[:, 1] = x[:, 1] * N0 + v * N1 + v * N2 + v * N3
    s1[:, 2] = x[:, 2] * N0 + x[:, 1] * N1 + v * N2 + v * N3
    s1[:, 3] = x[:, 3] * N0 + x[:, 2] * N1 + x[:, 1] * N2 + v * N3
    s1[:, 0] -= v * BN1 + v * BN2 + v * BN3 + v * BN4
    s1[:, 1] -= s1[:, 0] * D1 + v * BN2 + v * BN3 + v * BN4
    s1[:, 2] -= s1[:, 1] * D1 + s1[:, 0] * D2 + v * BN3 + v * BN4
    s1[:, 3] -= s1[:, 2] * D1 + s1[:, 1] * D2 + s1[:, 0] * D3 + v * BN4
    for i in range(4, ln):
        s1[:, i] = x[:, i] * N0 + x[:, i - 1] * N1 + x[:, i - 2] * N2 + x[:, i - 3] * N3
        s1[:, i] -= s1[:, i - 1] * D1 + s1[:, i - 2] * D2 + s1[:, i - 3] * D3 + s1[:, i - 4] * D4
    s2 = np.zeros_like(x)
    w = x[:, ln - 1]
    s2[:, ln - 1] = w * M1 + w * M2 + w * M3 + w * M4
    s2[:, ln - 2] = x[:, ln - 1] * M1 + w * M2 + w * M3 + w * M4
    s2[:, ln - 3] = x[:, ln - 2] * M1 + x[:, ln - 1] * M2 + w * M3 + w * M4
    s2[:, ln - 4] = x[:, ln - 3] * M1 + x[:, ln - 2] * M2 + x[:, ln - 1] * M3 + w * M4
    s2[:, ln - 1] -= w * BM1 + w * BM2 + w * BM3 + w * BM4
    s2[:, ln - 2] -= s2[:, ln - 1] * D1 + w * BM2 + w * BM3 + w * BM4
    s2[:, ln - 3] -= s2[:, ln - 2] * D1 + s2[:, ln - 1] * D2 + w * BM3 + w * BM4
    s2[:, ln - 4] -= s2[:, ln - 3] * D1 + s2[:, ln - 2] * D2 + s2[:, ln - 1] * D3 + w * BM4
    for i in range(ln - 4, 0, -1):
        s2[:, i - 1] = x[:, i] * M1 + x[:, i + 1] * M2 + x[:, i + 2] * M3 + x[:, i + 3] * M4
        s2[:, i - 1] -= s2[:, i] * D1 + s2[:, i + 1] * D2 + s2[:, i + 2] * D3 + s2[:, i + 3] * D4
    return s1 + s2


def _filter_axis(arr, axis, coef):
    """The recursion along ``axis`` of a 2-D array; float32 output pixels (ITK's internal image type), as float64."""
    a = np.asarray(arr, dtype=np.float64)
    lines = a if axis == 1 else a.T
    out = recursive_gaussian_lines(np.ascontiguousarray(lines), coef)
    out = out if axis == 1 else out.T
    return out.astype(np.float32).astype(np.float64)


def log_image(image, sigma):
    """pyradiomics ``getLoGImage`` -> ``sitk.LaplacianRecursiveGaussianImageFilter`` (NormalizeAcrossScale True,
    spacing (1, 1)): for each dimension the second-order recursive Gaussian along it and the zero-order one along
    the other, intermediate images float32, summed in float32.  Returned as float64 holding the float32 values."""
    a = np.asarray(image).astype(np.float64)
    if min(a.shape) < 4:
        raise ValueError("image too small for the recursive Gaussian (needs >= 4 pixels per axis)")
    c0 = _deriche_coefficients(float(sigma), 0)
    c2 = _deriche_coefficients(float(sigma), 2)
    total = np.zeros(a.shape, dtype=np.float32)
    for itk_dim in (0, 1):                  # ITK dimension 0 = x = NumPy axis 1
        ax = 1 - itk_dim
        d = _filter_axis(a, ax, c2)         # derivative filter first (its input is the image itself)
        s = _filter_axis(d, 1 - ax, c0)     # then the smoothing filter along the other dimension
        total = (total + s.astype(np.float32)).astype(np.float32)
    return total.astype(np.float64)


def log_name(sigma):
    """pyradiomics: ``'log-sigma-%s-mm-3D' % str(sigma).replace('.', '-')`` (params.yml sigma [1.0, 2.0, 3.0])."""
    return "log-sigma-%s-mm-3D" % str(sigma).replace(".", "-")


def filtered_images(image, image_type, options=None, force2D=False, force2Ddimension=0):
    """``OrderedDict[name -> array]`` of the images one enabled imageType yields, in pyradiomics' order."""
    options = options or {}
    if image_type == "Gradient":
        return OrderedDict([("gradient", gradient_image(image))])
    if image_type == "LoG":
        return OrderedDict((log_name(s), log_image(image, s)) for s in options.get("sigma", []))
    if image_type == "Wavelet":
        return wavelet_images(image, force2D, force2Ddimension)
    raise ValueError("not a filtered image type: %r" % image_type)
