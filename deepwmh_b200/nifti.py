"""Minimal NIfTI-1 single-file reader / writer (.nii, .nii.gz).

The reference reads and writes NIfTI through nibabel (deepwmh/utilities/data_io.py:223-263,285-286) and
nnU-Net through SimpleITK; neither is a dependency here.  Arrays are returned as nibabel's `get_fdata()`
returns them: indexed [x, y, z] with the file's x fastest, intensity scaling (scl_slope/scl_inter) applied.
Only what the prediction path needs is implemented: 3-D (or 4-D with one volume) images of the common
scalar dtypes, header preserved byte-for-byte on write except for datatype / bitpix / scaling / dim.
"""
from __future__ import annotations

import gzip
import struct
from typing import Dict, Tuple

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64,
           256: np.int8, 512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def read_nifti(path) -> Tuple[np.ndarray, Dict]:
    """-> (data float32 [x,y,z], header dict).  header['raw'] holds the 348 header bytes (+ extension bytes)."""
    with _open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 348:
        raise ValueError("%s: not a NIfTI-1 file (too short)" % path)
    endian = "<"
    if struct.unpack("<i", buf[0:4])[0] != 348:
        if struct.unpack(">i", buf[0:4])[0] != 348:
            raise ValueError("%s: sizeof_hdr != 348 (NIfTI-2 / Analyze are not supported)" % path)
        endian = ">"
    magic = buf[344:348]
    if magic[:3] not in (b"n+1", b"ni1"):
        raise ValueError("%s: bad NIfTI magic %r" % (path, magic))
    if magic[:3] == b"ni1":
        raise ValueError("%s: two-file NIfTI (.hdr/.img) is not supported" % path)
    dim = struct.unpack(endian + "8h", buf[40:56])
    datatype, bitpix = struct.unpack(endian + "hh", buf[70:74])
    pixdim = struct.unpack(endian + "8f", buf[76:108])
    vox_offset = int(struct.unpack(endian + "f", buf[108:112])[0])
    slope, inter = struct.unpack(endian + "ff", buf[112:120])
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, datatype))
    nd = dim[0]
    shape = [int(d) for d in dim[1:1 + nd]]
    while len(shape) > 3 and shape[-1] == 1:
        shape.pop()
    if len(shape) != 3:
        raise ValueError("%s: expected a single 3-D volume, got dim %s" % (path, dim[:nd + 1]))
    n = int(np.prod(shape))
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    arr = np.frombuffer(buf, dtype=dt, count=n, offset=vox_offset).reshape(shape, order="F")
    data = arr.astype(np.float32)
    if slope not in (0.0, 1.0) or (inter != 0.0 and slope != 0.0):
        if slope != 0.0 and np.isfinite(slope):
            data = data * np.float32(slope) + np.float32(inter)
    qform_code, sform_code = struct.unpack(endian + "hh", buf[252:256])
    hdr = {
        "raw": bytes(buf[:vox_offset]), "endian": endian, "dim": dim, "pixdim": pixdim, "datatype": datatype,
        "spacing": tuple(float(abs(p)) for p in pixdim[1:4]), "qform_code": qform_code, "sform_code": sform_code,
        "srow": np.array(struct.unpack(endian + "12f", buf[280:328]), dtype=np.float64).reshape(3, 4),
        "shape": tuple(shape),
    }
    return np.ascontiguousarray(data), hdr


def default_header(shape, spacing=(1.0, 1.0, 1.0)) -> Dict:
    """Identity-orientation header (what data_io.save_nifti_simple produces via nibabel with affine = eye)."""
    raw = bytearray(352)
    struct.pack_into("<i", raw, 0, 348)
    struct.pack_into("<8h", raw, 40, 3, shape[0], shape[1], shape[2], 1, 1, 1, 1)
    struct.pack_into("<hh", raw, 70, 16, 32)
    struct.pack_into("<8f", raw, 76, 1.0, spacing[0], spacing[1], spacing[2], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", raw, 108, 352.0)
    struct.pack_into("<ff", raw, 112, 1.0, 0.0)
    raw[123] = 2                                                    # xyzt_units: mm
    struct.pack_into("<hh", raw, 252, 0, 2)                          # sform_code = aligned
    struct.pack_into("<12f", raw, 280, spacing[0], 0, 0, 0, 0, spacing[1], 0, 0, 0, 0, spacing[2], 0)
    raw[344:348] = b"n+1\x00"
    return {"raw": bytes(raw), "endian": "<", "spacing": tuple(spacing), "shape": tuple(shape)}


def write_nifti(path, data: np.ndarray, header: Dict, dtype=np.float32):
    """Write `data` [x,y,z] with the geometry of `header` (as returned by read_nifti / default_header)."""
    data = np.asarray(data)
    if data.ndim != 3:
        raise ValueError("write_nifti expects a 3-D array")
    dt = np.dtype(dtype)
    if dt.str[1:] not in _CODES:
        raise ValueError("unsupported dtype %s" % dt)
    endian = header.get("endian", "<")
    raw = bytearray(header["raw"])
    if len(raw) < 352:
        raw.extend(b"\x00" * (352 - len(raw)))
    vox_offset = len(raw)
    struct.pack_into(endian + "8h", raw, 40, 3, data.shape[0], data.shape[1], data.shape[2], 1, 1, 1, 1)
    struct.pack_into(endian + "hh", raw, 70, _CODES[dt.str[1:]], dt.itemsize * 8)
    struct.pack_into(endian + "f", raw, 108, float(vox_offset))
    struct.pack_into(endian + "ff", raw, 112, 1.0, 0.0)
    payload = np.asarray(data, dtype=dt.newbyteorder(endian)).tobytes(order="F")
    with _open(path, "wb") as f:
        f.write(bytes(raw))
        f.write(payload)
