"""Example 6: recovering the rotation of a loaded mesh (the reference's teapot) from a target image
(src/examples/example6.py).  Converges for small perturbations, as the reference notes."""
import argparse

import numpy as np
import torch

from .. import camera_utils, obj_utils, shapes
from ..render import render
from . import image_io

IMAGE_WIDTH, IMAGE_HEIGHT = 640, 480


def render_with_rotation(euler_angles, mesh, device):
    """example6.py:40-63: camera at (0,3,3) looking at the origin, one light above."""
    vertices, triangles, normals = mesh
    model_rotation = camera_utils.euler_matrices(euler_angles)[0, :3, :3]
    vertices_world_space = torch.matmul(vertices, model_rotation.T)
    # normals transform with the inverse transpose of the model matrix (example6.py:56-57)
    normals_world_space = torch.matmul(normals, torch.inverse(model_rotation.T).T)
    eye = torch.tensor([[0.0, 3.0, 3.0]], dtype=torch.float32, device=device)
    center = torch.tensor([[0.0, 0.0, 0.0]], dtype=torch.float32, device=device)
    world_up = torch.tensor([0.0, np.cos(-np.pi / 4.0), np.sin(-np.pi / 4.0)], dtype=torch.float32, device=device)
    vertex_diffuse_colors = torch.ones_like(vertices, dtype=torch.float32)
    light_positions = torch.tensor([[[0.0, 3.0, 0.0]]], dtype=torch.float32, device=device)
    light_intensities = torch.ones([1, 1, 3], dtype=torch.float32, device=device)
    image = render(vertices_world_space, triangles, normals_world_space, vertex_diffuse_colors, eye, center,
                   world_up, light_positions, light_intensities, IMAGE_WIDTH, IMAGE_HEIGHT)
    return torch.reshape(image, [IMAGE_HEIGHT, IMAGE_WIDTH, 4])


def fit_mesh_rotation(mesh, target_render, initial_euler_angles, epochs=50, writer=None, device=None, log=None):
    """mesh = (vertices [V,3], triangles [T,3], normals [V,3]) -> (euler angles [1,3], losses)."""
    device = torch.device(device or "cuda")
    vertices, triangles, normals = mesh
    mesh = (vertices.to(device)[None, :, :], triangles.to(device), normals.to(device)[None, :, :])
    target_render = target_render.to(device)
    euler_angles = torch.tensor(initial_euler_angles, dtype=torch.float32, device=device, requires_grad=True)
    optimizer = torch.optim.SGD([euler_angles], 0.7, 0.1)

    def stepfn():
        optimizer.zero_grad()
        image = render_with_rotation(euler_angles, mesh, device)
        if writer is not None:
            writer.append_data(image_io.frame_on_black(image.detach().cpu().numpy()))
        loss = torch.mean(torch.abs(image - target_render))
        loss.backward()
        torch.nn.utils.clip_grad_norm_([euler_angles], 1.0)
        return loss

    losses = []
    for e in range(epochs):
        loss = optimizer.step(stepfn)
        losses.append(float(loss.detach()))
        if log is not None:
            log("step {} of {}: loss {:.5f}".format(e, epochs, losses[-1]))
    return euler_angles.detach(), losses


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__)
    parser.add_argument("-i", "--filename_input", type=str, default=None,
                        help=".obj file (the reference ships teapot.obj; without one a flattened sphere is used)")
    parser.add_argument("-t", "--filename_target", type=str, default=None,
                        help="RGBA target image 640x480; without one the mesh rendered unrotated is the target")
    parser.add_argument("-o", "--filename_output", type=str, default="example6.gif")
    args = parser.parse_args(argv)
    device = torch.device("cuda")
    if args.filename_input:
        mesh = obj_utils.load_obj(args.filename_input)
    else:
        vertices, triangles, normals = shapes.sphere(1.0, resolution=25)
        scale = torch.tensor([1.0, 0.4, 0.7])
        mesh = (vertices * scale, torch.flip(triangles, [1]), torch.nn.functional.normalize(normals / scale, dim=-1))
    if args.filename_target:
        target = torch.tensor(image_io.imread(args.filename_target).astype(float) / 255.0)
    else:
        on_device = (mesh[0].to(device)[None], mesh[1].to(device), mesh[2].to(device)[None])
        target = render_with_rotation(torch.zeros(1, 3, device=device), on_device, device).detach()
    writer = image_io.FrameWriter(args.filename_output, fps=20)
    angles, losses = fit_mesh_rotation(mesh, target, [[np.pi / 4.0, 0.0, 0.0]], writer=writer, device=device,
                                       log=print)
    writer.close()
    print("euler angles:", angles.cpu().tolist(), "final loss:", losses[-1])


if __name__ == "__main__":
    main()
