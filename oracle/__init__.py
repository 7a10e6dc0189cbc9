"""CPU oracle for the rasterization hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  pytorch_mesh_renderer_b200 never does.
"""
