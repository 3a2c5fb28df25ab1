"""Default parameters of the energy-grid path — the VALUES are part of parity with the reference
(gauNEGF/config.py:8-33; the README table there is stale, the code wins: SURVEY.md appendix A.1)."""

# Physical parameters
TEMPERATURE = 0.0               # K
ETA = 1e-6                      # eV broadening
ENERGY_STEP = 0.001             # eV

# Contact tolerances
FERMI_CALCULATION_TOL = 1e-3
FERMI_SEARCH_CYCLES = 10
SURFACE_GREEN_CONVERGENCE = 1e-5
SURFACE_RELAXATION_FACTOR = 0.1

# Integration parameters
ADAPTIVE_INTEGRATION_TOL = 1e-4
N_KT = 10
ENERGY_MIN = -1e6
MAX_CYCLES = 1000
MAX_GRID_POINTS = 1000

# Fixed-point iteration caps hard-coded in the reference (surfG1D.py:265, surfGBethe.py:998,1076)
SURFACE_GREEN_MAX_ITER = 2000
BETHE_MAX_ITER = 1000
BETHE_MIXING = 0.5

# Logging: the reference opens a log file in CWD at import (integrate.py:28-45); this build only
# logs through logging.getLogger('gauNEGF.integrate') and never creates files on import.
LOG_LEVEL = 'DEBUG'
LOG_PERFORMANCE = False
