"""ncu target for the kernels around the stage kernels: one resident-impulse frame (f2), one colouring (f3), one PredictBoundingBoxes (f4) on C2."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bepuphysics2_b200 as bp
from bepuphysics2_b200 import coloring, native, scenes

scene = scenes.shape_pile(100_000, seed=5)
sim = bp.Simulation(substeps=2, velocity_iterations=1)
scenes.build(scene, sim)
ts = bp.CudaTimestepper(sim)
ts.register_host_buffers()
pool, _ = ts.contact_feature_pool(np.random.default_rng(11))
ts.register_array(pool)
ts.describe()
ts.set_contact_feature_pool(pool)
for _ in range(2):
    ts.update_contacts_from_pool(pool)
    ts.solve_device_only(1 / 60)
    ts.download_body_motion()
refs = coloring.scene_references(scene)
ts.color_constraints(refs, sim.body_count, 64, order=1)
shapes = np.zeros(sim.body_count, dtype=native.BODY_SHAPE_DTYPE)
shapes["type"] = np.random.default_rng(3).choice([0, 1, 2, 4], size=sim.body_count)
shapes["a"] = shapes["b"] = shapes["c"] = 0.5
shapes["maximum_speculative_margin"] = 3.4e38
activities = np.zeros(sim.body_count, dtype=native.BODY_ACTIVITY_DTYPE)
ts.set_body_shapes(shapes)
ts.predict_bounding_boxes(1 / 60, activities)
ts.synchronize()
ts.close()
print("done")
