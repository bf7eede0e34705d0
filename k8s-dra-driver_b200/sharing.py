"""Host-side sharing-config arithmetic that feeds ``ClaimRec.mem_limit_mib`` (spec §7).

Mirrors, name for name, the part of the reference's opaque-config API that turns MPS pinned-memory limits
into per-device ``"<N>M"`` strings:

* ``MpsPerDevicePinnedMemoryLimit.Normalize``  api/nvidia.com/resource/gpu/v1alpha1/sharing.go:190-209
* ``limit.get`` / ``limit.Megabyte``           sharing.go:213-237
* ``uuidSet.Normalize``                        sharing.go:257-273
* ``TimeSliceInterval.Int``                    sharing.go:168-180
* ``MpsConfig.Validate``                       validate.go:56-66

and the subset of ``k8s.io/apimachinery/pkg/api/resource.Quantity`` parsing those functions rely on
(``MustParse`` + ``Value()``: decimal/binary SI suffixes, value rounded up away from zero).
The golden vectors of sharing_test.go:37-149 are committed under tests/golden/mps_limits.json.
"""
from __future__ import annotations

import re
from fractions import Fraction

_BIN = {"Ki": 2 ** 10, "Mi": 2 ** 20, "Gi": 2 ** 30, "Ti": 2 ** 40, "Pi": 2 ** 50, "Ei": 2 ** 60}
_DEC = {"n": Fraction(1, 10 ** 9), "u": Fraction(1, 10 ** 6), "m": Fraction(1, 10 ** 3), "": Fraction(1),
        "k": Fraction(10 ** 3), "M": Fraction(10 ** 6), "G": Fraction(10 ** 9), "T": Fraction(10 ** 12),
        "P": Fraction(10 ** 15), "E": Fraction(10 ** 18)}
_Q = re.compile(r"^([+-]?)(\d+(?:\.\d*)?|\.\d+)(Ki|Mi|Gi|Ti|Pi|Ei|[numkMGTPE]|[eE][+-]?\d+)?$")


class ErrInvalidDeviceSelector(ValueError):
    """sharing.go:183 — a device index or UUID was invalid."""


class ErrInvalidLimit(ValueError):
    """sharing.go:186 — a limit was invalid."""


def quantity_value(q) -> int:
    """resource.MustParse(q).Value(): integer, rounded up away from zero."""
    if isinstance(q, int):
        return q
    m = _Q.match(q.strip()) if isinstance(q, str) else None
    if not m:
        raise ValueError(f"quantities must match the regular expression: {q!r}")
    sign, num, suf = m.group(1), m.group(2), m.group(3) or ""
    v = Fraction(num)
    if suf in _BIN:
        v *= _BIN[suf]
    elif suf in _DEC:
        v *= _DEC[suf]
    else:
        v *= Fraction(10) ** int(suf[1:])
    n = -(-v.numerator // v.denominator)       # ceil for the magnitude
    return -n if sign == "-" else n


def megabyte(q):
    """limit.Megabyte (sharing.go:234-237): (\"<v>M\", v > 0) with v = Value()/1024/1024 truncated."""
    val = quantity_value(q)
    v = abs(val) // 1024 // 1024
    if val < 0:
        v = -v
    return f"{v}M", v > 0


def megabyte_mib(q) -> int:
    """The integer ClaimRec.mem_limit_mib carries; raises ErrInvalidLimit when Megabyte() says invalid."""
    s, ok = megabyte(q)
    if not ok:
        raise ErrInvalidLimit(f"value set too low: {q}")
    return int(s[:-1])


def _uuid_normalize(uuids, key: str) -> str:
    """uuidSet.Normalize (sharing.go:257-273)."""
    if key in set(uuids):
        return key
    if not re.match(r"^[+-]?\d+$", key):       # strconv.Atoi
        raise ErrInvalidDeviceSelector(f"unable to parse key as an integer: {key}")
    index = int(key)
    if 0 <= index < len(uuids):
        return uuids[index]
    raise ErrInvalidDeviceSelector(f"invalid device index: {index}")


def normalize(per_device_limit: dict | None, uuids, default_limit=None) -> dict:
    """MpsPerDevicePinnedMemoryLimit.Normalize(uuids, defaultPinnedDeviceMemoryLimit) (sharing.go:190-209)."""
    uuids = list(uuids or [])
    limits: dict = {}
    if default_limit is not None and uuids:    # limit.get, sharing.go:213-228
        s, ok = megabyte(default_limit)
        if not ok:
            raise ErrInvalidLimit(f"default value set too low: {default_limit}")
        for u in uuids:
            limits[u] = s
    for k, v in (per_device_limit or {}).items():
        dev = _uuid_normalize(uuids, k)
        s, ok = megabyte(v)
        if not ok:
            raise ErrInvalidLimit(f"value set too low: {k}: {v}")
        limits[dev] = s
    return limits


_TS = {"Default": 0, "Short": 1, "Medium": 2, "Long": 3}


def time_slice_int(interval: str) -> int:
    """TimeSliceInterval.Int (sharing.go:168-180)."""
    return _TS.get(interval, -1)


def validate_mps(default_active_thread_percentage=None) -> None:
    """MpsConfig.Validate (validate.go:56-66): thread percentage within [0, 100]."""
    p = default_active_thread_percentage
    if p is not None and (p < 0 or p > 100):
        raise ValueError("active thread percentage must be in [0, 100]")
