"""NumPy restatement of gaussian_deconvolve (pyGPA/geometric_phase_analysis.py:892-904) and of the
third-party routine it rests on (oracle; test infrastructure only).

PARITY UNPINNED for this module: the arithmetic lives in scikit-image's
``skimage.restoration.wiener`` / ``skimage.restoration.uft`` (requirements.txt of the reference
lists scikit-image with no version pin; the package is absent from this image and there is no
network), so the reference function cannot be executed here and no golden vector exists.  What
follows restates the published algorithm (skimage/restoration/deconvolution.py, uft.py, as of
scikit-image 0.19-0.22):
    reg        = uft.laplacian(2, shape)  -> ir2tf of the 3x3 stencil [[0,-1,0],[-1,4,-1],[0,-1,0]]
    trans_func = uft.ir2tf(psf, shape)    -> rfftn of the PSF rolled so that its centre sits at [0,0]
    filter     = conj(trans_func) / (|trans_func|^2 + balance |reg|^2)
    deconv     = uirfft2(filter * urfft2(image))          (unitary transforms: the norms cancel)
and anchors parity on the reference's call site (gpa.py:892-904: reflect padding by 2 dr, the PSF
built from scipy.ndimage.fourier_gaussian, balance = 5000, clip=False, crop)."""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi

__all__ = ["gaussian_deconvolve", "wiener_transfer"]


def _ir2tf(imp_resp, shape):
    """skimage.restoration.uft.ir2tf(imp_resp, shape, is_real=True) for 2-D data."""
    irpadded = np.zeros(shape)
    irpadded[tuple(slice(0, s) for s in imp_resp.shape)] = imp_resp
    for axis, axis_size in enumerate(imp_resp.shape):
        irpadded = np.roll(irpadded, shift=-int(np.floor(axis_size / 2)), axis=axis)
    return np.fft.rfftn(irpadded, axes=(-2, -1))


def _laplacian_tf(shape):
    """skimage.restoration.uft.laplacian(2, shape, is_real=True)[0]."""
    impr = np.zeros((3, 3))
    impr[1, :] = -1.0
    impr[:, 1] = -1.0
    impr[1, 1] = 4.0
    return _ir2tf(impr, shape)


def _wiener(image, psf, balance):
    """skimage.restoration.wiener(image, psf, balance, clip=False, is_real=True)."""
    reg = _laplacian_tf(image.shape)
    trans_func = _ir2tf(psf, image.shape) if psf.shape != reg.shape else psf
    wiener_filter = np.conj(trans_func) / (np.abs(trans_func) ** 2 + balance * np.abs(reg) ** 2)
    return np.fft.irfft2(wiener_filter * np.fft.rfft2(image), s=image.shape)


def gaussian_deconvolve(data, sigma, dr=20, balance=5000):
    """geometric_phase_analysis.py:892-904."""
    data = np.asarray(data, dtype=np.float64)
    padding = [(0, 0)] * (data.ndim - 2) + [(2 * dr, 2 * dr), (2 * dr, 2 * dr)]
    padded = np.pad(data, padding, mode='reflect')
    kernel = np.fft.fft2(ndi.fourier_gaussian(np.ones(padded.shape[-2:]), sigma=sigma)).real
    kernel = np.fft.fftshift(kernel)
    kernel = kernel / kernel.sum()
    planes = [_wiener(p, kernel, balance)[2 * dr:-2 * dr, 2 * dr:-2 * dr]
              for p in padded.reshape((-1,) + padded.shape[-2:])]
    return np.reshape(np.stack(planes), data.shape)


def wiener_transfer(shape, sigma, balance=5000):
    """The closed form the CUDA kernel evaluates: for the PSF above trans_func is the Gaussian
    transfer function itself and reg is 4 - 2 cos(2 pi fx) - 2 cos(2 pi fy) (full-spectrum layout)."""
    fx = np.fft.fftfreq(shape[0])[:, None]
    fy = np.fft.fftfreq(shape[1])[None, :]
    h = np.exp(-2 * np.pi ** 2 * sigma ** 2 * (fx ** 2 + fy ** 2))
    lap = 4 - 2 * np.cos(2 * np.pi * fx) - 2 * np.cos(2 * np.pi * fy)
    return h / (h ** 2 + balance * lap ** 2)
