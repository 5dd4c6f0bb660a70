"""Python twin of the pose generator of dump_reference.jl (SplitMix64 + Box-Muller, seed 1234): regenerates the poses
the Julia dump used, as a cross-check that a julia_reference.json on disk belongs to this script version."""
import math

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed=1234):
        self.s = seed & MASK

    def u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def uniform(self):
        return (float(self.u64() >> 11) + 0.5) / 9007199254740992.0

    def normal(self):
        u1 = self.uniform(); u2 = self.uniform()
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


def poses(g):
    """One (r1, q1, r2, q2) draw in the order dump_reference.jl makes it."""
    r1 = [2 * g.normal() for _ in range(3)]
    q1 = [g.normal() for _ in range(4)]; n = math.sqrt(sum(v * v for v in q1)); q1 = [v / n for v in q1]
    r2 = [g.normal() for _ in range(3)]
    q2 = [g.normal() for _ in range(4)]; n = math.sqrt(sum(v * v for v in q2)); q2 = [v / n for v in q2]
    return r1, q1, r2, q2
