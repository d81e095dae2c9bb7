"""The uniform tapes of the reference-pinning fixtures: splitmix64 → 53-bit uniforms, a dozen lines that will give the
same numbers forever (no dependency on a library's generator)."""
import numpy as np

_M = (1 << 64) - 1


def splitmix_tape(seed: int, n: int) -> np.ndarray:
    out = np.empty(n)
    s = seed & _M
    for k in range(n):
        s = (s + 0x9E3779B97F4A7C15) & _M
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
        z ^= z >> 31
        out[k] = (z >> 11) * 2.0 ** -53
    return out


def write_tape(path: str, tape) -> None:
    with open(path, "w") as f:
        f.write("\n".join(repr(float(x)) for x in tape) + "\n")
