"""Golden sweep of the reference's `space_timesteps` (respace.py:11-61), generated from the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_spacing.py

Writes spacing.json: for every spec a CRC32 of the sorted kept-step list (or the exception type the reference raises).
Specs: every single count 1..1000 over 1000 steps, every "ddimN" for N in 1..1000, a set of multi-section lists, and
the same families over 250 and 37 steps; plus `get_named_beta_schedule` (gaussian_diffusion.py:112-134) for both
schedules (and an unknown name) at eight lengths, as a CRC32 of the float64 bytes.  Build container only (the GPU box has no /root/reference).
"""
import json
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OSU_DIFFUSION_REF", "/root/reference")
sys.path.insert(0, REF)
from diffusion.respace import space_timesteps  # noqa: E402  (reference)
from diffusion.gaussian_diffusion import get_named_beta_schedule  # noqa: E402  (reference)


def specs():
    for n in (1000, 250, 37):
        for k in range(1, n + 1):
            yield n, str(k)
            yield n, f"ddim{k}"
        yield n, str(n + 1)
        for lst in ("10,10,10", "1,1,1,1", "5,0,5", "10,15,20", "3,7", "12,1,12,1", "33,33,34", "2,3,4,5,6,7",
                    "9,9,9,9,9,9,9,9,9,9", "1", "0", "36,1", "13,13,13"):
            yield n, lst


def digest(n, spec):
    try:
        kept = sorted(space_timesteps(n, spec))
    except Exception as e:  # the error behaviour is part of the surface
        return type(e).__name__
    return "%d:%08x" % (len(kept), zlib.crc32(",".join(map(str, kept)).encode()))


def main():
    out = {f"{n}|{spec}": digest(n, spec) for n, spec in specs()}
    for name in ("linear", "squaredcos_cap_v2", "cosine"):  # "betas|<name>|<n>": CRC32 of the float64 bytes
        for n in (1, 2, 10, 37, 100, 250, 1000, 4000):
            try:
                b = get_named_beta_schedule(name, n)
                out[f"betas|{name}|{n}"] = "%d:%08x" % (len(b), zlib.crc32(b.astype("<f8").tobytes()))
            except Exception as e:
                out[f"betas|{name}|{n}"] = type(e).__name__
    with open(os.path.join(HERE, "spacing.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True)
    print(len(out), "specs;", sum(1 for v in out.values() if ":" not in v), "raise")


if __name__ == "__main__":
    main()
