"""Image IO of the examples with PIL only: the three calls the reference's examples make on skimage.io and
imageio (imread, imsave, get_writer/append_data/close)."""
import numpy as np
from PIL import Image


def imread(path):
    """-> uint8 array [H,W,C] (skimage.io.imread for the PNG fixtures)."""
    return np.asarray(Image.open(path))


def imsave(path, image):
    Image.fromarray(np.ascontiguousarray(image, dtype=np.uint8)).save(path)


def to_uint8(image):
    """float image in [0,1] -> uint8, as the examples do before writing (example1.py:49-51)."""
    return (np.clip(image, 0.0, 1.0) * 255.0).astype(np.uint8)


def frame_on_black(render):
    """RGBA float render -> RGB premultiplied over black with opaque alpha (example5.py:72-77)."""
    frame = np.concatenate([render[:, :, :3] * render[:, :, 3][:, :, None],
                            np.ones([render.shape[0], render.shape[1], 1], dtype=np.float32)], axis=-1)
    return to_uint8(frame)


class FrameWriter:
    """Collects frames and writes an animated GIF (the reference writes .mp4 through imageio/ffmpeg); a path
    ending in .png writes numbered stills instead.  `path=None` collects nothing."""

    def __init__(self, path, fps=20):
        self.path, self.fps, self.frames = path, fps, []

    def append_data(self, frame):
        if self.path is not None:
            self.frames.append(Image.fromarray(np.ascontiguousarray(frame[:, :, :3], dtype=np.uint8)))

    def close(self):
        if self.path is None or not self.frames:
            return
        if self.path.lower().endswith(".png"):
            stem = self.path[:-4]
            for k, frame in enumerate(self.frames):
                frame.save("%s_%04d.png" % (stem, k))
        else:
            self.frames[0].save(self.path, save_all=True, append_images=self.frames[1:],
                                duration=int(1000 / self.fps), loop=0)
        self.frames = []
