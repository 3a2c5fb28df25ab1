"""Energy-independent self-energy provider — drop-in for gauNEGF/surfGTester.py (surfG protocol:
sigma(E, i), sigmaTot(E), setF(F, mu1, mu2); attributes F, S)."""
import numpy as np

from .config import SURFACE_GREEN_CONVERGENCE
from .matTools import formSigma


class surfGTest:
    def __init__(self, Fock, Overlap, indsList, sig1=None, sig2=None):
        self.F = Fock
        self.S = Overlap
        self.N = len(Fock)
        self.indsList = indsList
        if sig1 is None:
            # the reference's default branch cannot run (surfGTester.py:83,91-92: an N-sized diagonal
            # assigned into a contact-sized block); give the documented default -0.05j instead
            sig1 = -0.05j
        self.sig = [formSigma(indsList[0], sig1, self.N, self.S),
                    formSigma(indsList[1], sig1 if sig2 is None else sig2, self.N, self.S)]

    def sigma(self, E, i, conv=SURFACE_GREEN_CONVERGENCE):
        return self.sig[i]

    def sigmaTot(self, E, conv=SURFACE_GREEN_CONVERGENCE):
        sigTot = np.array(np.zeros((self.N, self.N)), dtype=complex)
        for i in range(len(self.indsList)):
            sigTot += self.sigma(E, i, conv)
        return sigTot

    def setF(self, F, mu1=None, mu2=None):
        self.F = F
