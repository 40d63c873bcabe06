"""Wall shear stress of a stitched prediction (reference compute_wss.py), computed on the GPU.

    from compute_wss import compute_wall_shear_stress
    surface, wss, wss_magnitude = compute_wall_shear_stress(stitched, 'velocity', dynamic_viscosity=1.0e-3)
"""
from fesr_b200.postprocess import compute_wall_shear_stress  # noqa: F401
