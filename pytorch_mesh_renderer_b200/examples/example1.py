"""Example 1: rendering a mesh from a fixed camera (src/examples/example1.py): load an .obj, white diffuse
colour, one light above, 640x480."""
import argparse

import torch

from .. import obj_utils, shapes
from ..render import render
from . import image_io


def render_obj(vertices, triangles, normals, image_width=640, image_height=480, device=None):
    """vertices / normals [V,3], triangles [T,3] -> RGBA float image [H,W,4] on `device` (example1.py:25-48)."""
    device = torch.device(device or "cuda")
    vertices = vertices.to(device)[None, :, :]
    normals = normals.to(device)[None, :, :]
    eye = torch.tensor([[0.0, 0.0, 3.0]], dtype=torch.float32, device=device)
    center = torch.tensor([[0.0, 0.0, 0.0]], dtype=torch.float32, device=device)
    world_up = torch.tensor([[0.0, 1.0, 0.0]], dtype=torch.float32, device=device)
    vertex_diffuse_colors = torch.ones_like(vertices, dtype=torch.float32)
    light_positions = torch.tensor([[[0.0, 3.0, 0.0]]], dtype=torch.float32, device=device)
    light_intensities = torch.ones([1, 1, 3], dtype=torch.float32, device=device)
    image = render(vertices, triangles.to(device), normals, vertex_diffuse_colors, eye, center, world_up,
                   light_positions, light_intensities, image_width, image_height)
    return torch.reshape(image, [image_height, image_width, 4])


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__)
    parser.add_argument("-i", "--filename_input", type=str, default=None,
                        help=".obj file (the reference ships teapot.obj; without one a sphere is rendered)")
    parser.add_argument("-o", "--filename_output", type=str, default="example1.png")
    args = parser.parse_args(argv)
    if args.filename_input:
        vertices, triangles, normals = obj_utils.load_obj(args.filename_input)
    else:
        vertices, triangles, normals = shapes.sphere(1.0, resolution=25)
        triangles = torch.flip(triangles, [1])          # the generators wind CCW, the renderer's cube tests CW
    image = render_obj(vertices, triangles, normals)
    image_io.imsave(args.filename_output, image_io.to_uint8(image.cpu().numpy()))


if __name__ == "__main__":
    main()
