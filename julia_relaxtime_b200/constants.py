"""Model constants of the PNJL path, in fm units — mirror of src/Constants_PNJL.jl:82-103 of the reference,
with the values of config/pnjl/default.toml (identical to the built-in defaults at Constants_PNJL.jl:20-43).

A different parameter profile is passed by constructing `PNJLConstants(...)` with other MeV-level inputs,
the way the reference's TOML profiles do.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class PNJLConstants:
    hbarc: float = 197.327          # MeV fm
    N_color: int = 3
    rho0_fm3: float = 0.16          # fm^-3
    Lambda_MeV: float = 602.3
    G_over_Lambda2: float = 1.835
    K_over_Lambda5: float = 12.36
    m_ud0_MeV: float = 5.5
    m_s0_MeV: float = 140.7
    T0_MeV: float = 210.0
    a0: float = 3.51
    a1: float = -2.47
    a2: float = 15.2
    b3: float = -1.75

    # derived, fm units (Constants_PNJL.jl:90-103)
    @property
    def Lambda_inv_fm(self):
        return self.Lambda_MeV / self.hbarc

    @property
    def m_ud0_inv_fm(self):
        return self.m_ud0_MeV / self.hbarc

    @property
    def m_s0_inv_fm(self):
        return self.m_s0_MeV / self.hbarc

    @property
    def G_fm2(self):
        L = self.Lambda_inv_fm
        return self.G_over_Lambda2 / (L * L)

    @property
    def K_fm5(self):
        return self.K_over_Lambda5 / self.Lambda_inv_fm ** 5

    @property
    def T0_inv_fm(self):
        return self.T0_MeV / self.hbarc


DEFAULT = PNJLConstants()
HBARC = DEFAULT.hbarc
