"""The reference's examples for the barycentric renderer (src/examples/example1.py, example5.py, example6.py)
on the CUDA path: same scenes, optimisers and command-line options; every tensor lives on the GPU and frames
are written with PIL (the reference needs skimage / imageio / matplotlib, none of which this package uses).

    python -m pytorch_mesh_renderer_b200.examples.example1 -i mesh.obj -o example1.png
    python -m pytorch_mesh_renderer_b200.examples.example5 -o example5.gif
    python -m pytorch_mesh_renderer_b200.examples.example6 -i mesh.obj -o example6.gif

Each module exposes its scene as functions (`render_obj`, `fit_cube_rotation`, `fit_mesh_rotation`) so that the
tests drive exactly what the command line runs.
"""
