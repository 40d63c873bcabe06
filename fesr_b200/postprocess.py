"""Post-processing of the stitched field with the reference's interface (compute_wss.py:5-120), on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def compute_wall_shear_stress(grid, velocity_array_name='velocity', wall_boundary_ids=None, dynamic_viscosity=1.0,
                              output_filename=None):
    """Same arguments and return triple as the reference's compute_wall_shear_stress:
    (surface, wall_shear_stress [M, 3], wall_shear_stress_magnitude [M]) as numpy arrays, `surface` being a dict with
    the surface point ids, the outward-oriented boundary triangles and the point normals (the reference returns a
    vtkPolyData carrying the same arrays).  `grid`: a StitchedMesh (reconstruct_from_partition's result) or anything
    with `.pos` [N, 3], `.cells` [C, 4] and point arrays under `.point_data` or as attributes ('velocity' of a
    StitchedMesh is the stitched prediction, 'ref_velocity' the stitched reference).  `wall_boundary_ids` is accepted
    and unused, as in the reference.  `output_filename` (optional) gets an .npz with the result."""
    dev = torch.device("cuda", torch.cuda.current_device())
    pos = torch.as_tensor(np.asarray(grid.pos), dtype=torch.float32).to(dev)
    cells = torch.as_tensor(np.asarray(grid.cells), dtype=torch.int32).to(dev)
    vel = None
    if velocity_array_name == 'velocity' and hasattr(grid, "field"):
        vel = torch.as_tensor(grid.field)[:, :3]
    elif velocity_array_name == 'ref_velocity' and hasattr(grid, "ref_field"):
        vel = torch.as_tensor(grid.ref_field)[:, :3]
    else:
        pd = getattr(grid, "point_data", None)
        arr = pd.get(velocity_array_name) if isinstance(pd, dict) else getattr(grid, velocity_array_name, None)
        if arr is not None:
            vel = torch.as_tensor(np.asarray(arr))
    if vel is None:
        raise ValueError(f"Velocity array '{velocity_array_name}' not found in point data")
    if vel.shape[0] != pos.shape[0]:
        raise ValueError(f"velocity array has {vel.shape[0]} rows for {pos.shape[0]} points")
    out = ops.wall_shear_stress(pos, cells, vel.to(dev, dtype=torch.float32).contiguous(), dynamic_viscosity)
    wss = out["wss"].cpu().numpy()
    mag = out["wss_magnitude"].cpu().numpy()
    surface = {"point_ids": out["surface_nodes"].cpu().numpy(), "faces": out["faces"].cpu().numpy(),
               "Normals": out["normals"].cpu().numpy(), "WallShearStressVector": wss, "WallShearStressMagnitude": mag}
    print(f"Wall shear stress computed. Max magnitude: {mag.max() if mag.size else 0.0:.6f} Pa")
    print(f"Mean magnitude: {mag.mean() if mag.size else 0.0:.6f} Pa")
    if output_filename:
        np.savez(output_filename, **surface)
    return surface, wss, mag
