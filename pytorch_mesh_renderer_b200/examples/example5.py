"""Example 5: recovering the rotation of a cube from one target image by gradient descent through the
renderer (src/examples/example5.py): SGD(lr 0.7, momentum 0.1) on three Euler angles, L1 image loss,
gradient norm clipped to 1, 50 steps."""
import argparse

import torch

from .. import camera_utils, shapes
from ..render import render
from . import image_io

IMAGE_WIDTH, IMAGE_HEIGHT = 640, 480


def render_cube_with_rotation(euler_angles, cube, device):
    """example5.py:36-60: rotate the cube and its normals, camera at (0,0,6) carrying the light."""
    cube_vertices, cube_triangles, cube_normals = cube
    model_rotation = camera_utils.euler_matrices(euler_angles)[0, :3, :3]
    vertices_world_space = torch.reshape(torch.matmul(cube_vertices, model_rotation.T), [1, 8, 3])
    normals_world_space = torch.reshape(torch.matmul(cube_normals, model_rotation.T), [1, 8, 3])
    eye = torch.tensor([[0.0, 0.0, 6.0]], dtype=torch.float32, device=device)
    center = torch.tensor([[0.0, 0.0, 0.0]], dtype=torch.float32, device=device)
    world_up = torch.tensor([[0.0, 1.0, 0.0]], dtype=torch.float32, device=device)
    vertex_diffuse_colors = torch.ones_like(vertices_world_space, dtype=torch.float32)
    light_positions = torch.reshape(eye, [1, 1, 3])
    light_intensities = torch.ones([1, 1, 3], dtype=torch.float32, device=device)
    image = render(vertices_world_space, cube_triangles, normals_world_space, vertex_diffuse_colors, eye, center,
                   world_up, light_positions, light_intensities, IMAGE_WIDTH, IMAGE_HEIGHT)
    return torch.reshape(image, [IMAGE_HEIGHT, IMAGE_WIDTH, 4])


def fit_cube_rotation(target_render, epochs=50, writer=None, device=None, initial_euler_angles=((0.0, 0.0, 0.0),),
                      log=None):
    """-> (euler angles [1,3], losses).  target_render: float RGBA [480,640,4]."""
    device = torch.device(device or "cuda")
    vertices, triangles, normals = shapes.cube(2.0)
    cube = (vertices.to(device), torch.flip(triangles, [1]).to(device), normals.to(device))     # CCW -> CW
    target_render = target_render.to(device)
    euler_angles = torch.tensor(initial_euler_angles, dtype=torch.float32, device=device, requires_grad=True)
    optimizer = torch.optim.SGD([euler_angles], 0.7, 0.1)

    def stepfn():
        optimizer.zero_grad()
        image = render_cube_with_rotation(euler_angles, cube, device)
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
    parser.add_argument("-t", "--filename_target", type=str, default=None,
                        help="RGBA target image 640x480 (the reference uses test_data/Gray_Cube_0.png); without one "
                             "the cube rendered at Euler angles (-20, 0, 60) of mesh_renderer_test.py:36 is the target")
    parser.add_argument("-o", "--filename_output", type=str, default="example5.gif")
    args = parser.parse_args(argv)
    device = torch.device("cuda")
    if args.filename_target:
        target = torch.tensor(image_io.imread(args.filename_target).astype(float) / 255.0)
    else:
        vertices, triangles, normals = shapes.cube(2.0)
        cube = (vertices.to(device), torch.flip(triangles, [1]).to(device), normals.to(device))
        target = render_cube_with_rotation(torch.tensor([[-20.0, 0.0, 60.0]], device=device), cube, device).detach()
    writer = image_io.FrameWriter(args.filename_output, fps=20)
    angles, losses = fit_cube_rotation(target, writer=writer, device=device, log=print)
    writer.close()
    print("euler angles:", angles.cpu().tolist(), "final loss:", losses[-1])


if __name__ == "__main__":
    main()
