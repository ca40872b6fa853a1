"""TEST INFRASTRUCTURE — generates tests/golden/double_integrator_level_set.npz from the reference's data fixture.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_level_set
examples/data/time_optimal_control_for_double_integrator_results_from_level_set_methods.mat holds the minimum time to
reach the origin of the double integrator on a 101 x 101 (pos, vel) grid, from a level-set solver (``mttr``) and
analytically (``attr``); examples/double_integrator_optimal_time.ipynb cell 18 reads it, forms dV/dvel by central
differences over the velocity axis and steers with u = -sign(dV/dvel) at the nearest node.  The arrays below are the
notebook's own: ``value_level_set`` / ``value_analytic`` [vel, pos] (its ``.T``), the two axes (``np.linspace`` of the
file's grid description) and the velocity step of the central difference.  The fixture travels to the GPU box, where /root/reference
does not exist."""
import os

import numpy as np
import scipy.io

SRC = "/root/reference/examples/data/time_optimal_control_for_double_integrator_results_from_level_set_methods.mat"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "double_integrator_level_set.npz")


def main():
    m = scipy.io.loadmat(SRC)
    V = m["mttr"].T                                    # cell 18: value_function_by_level_set_method
    Va = m["attr"].T
    g = m["gridOut"]
    pos = np.linspace(g["min"][0][0][0][0], g["max"][0][0][0][0], g["N"][0][0][0][0])
    vel = np.linspace(g["min"][0][0][1][0], g["max"][0][0][1][0], g["N"][0][0][1][0])
    dv = g["dx"][0][0][1][0]
    np.savez_compressed(OUT, value_level_set=V, value_analytic=Va, pos=pos, vel=vel, dv=np.float64(dv))
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
