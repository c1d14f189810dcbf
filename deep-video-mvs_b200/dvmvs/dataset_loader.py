"""Image loading / pre-processing names the reference's test drivers import (dvmvs/dataset_loader.py:260-346):
load_image and PreprocessImage (crop, INTER_LINEAR resize, mean/std normalisation, intrinsics rescale).  The reference's
methods stay host-side (same results as the reference: they call the same cv2 functions); `load_image_u8` +
`PreprocessImage.apply_rgb_cuda` are the device path: upload the decoded uint8 frame once and
do crop + resize + colour order + normalisation + CHW in one kernel.  MVSDataset (training crawler) is out of scope."""
import cv2
import numpy as np


def load_image(path):
    return cv2.cvtColor(cv2.imread(str(path), cv2.IMREAD_COLOR).astype(np.float32), cv2.COLOR_BGR2RGB)


def load_image_u8(path):
    """The decoded frame as cv2 returns it ((H,W,3) uint8, BGR) -- input of PreprocessImage.apply_rgb_cuda."""
    return cv2.imread(str(path), cv2.IMREAD_COLOR)


class PreprocessImage:
    def __init__(self, K, old_width, old_height, new_width, new_height, distortion_crop=0, perform_crop=True):
        fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
        self.new_width, self.new_height, self.perform_crop = new_width, new_height, perform_crop
        self.crop_x = self.crop_y = 0
        width, height = float(old_width), float(old_height)
        if perform_crop:
            w_in, h_in = old_width - 2 * distortion_crop, old_height - 2 * distortion_crop
            target_ratio = float(new_width) / float(new_height)
            if float(w_in) / float(h_in) > target_ratio:          # too wide: crop columns
                self.crop_x = int(np.floor((w_in - h_in * target_ratio) / 2.0)) + distortion_crop
                self.crop_y = distortion_crop
            else:                                                 # too tall: crop rows
                self.crop_x = distortion_crop
                self.crop_y = int(np.floor((h_in - w_in / target_ratio) / 2.0)) + distortion_crop
            cx -= self.crop_x
            cy -= self.crop_y
            width, height = float(old_width - 2 * self.crop_x), float(old_height - 2 * self.crop_y)
        sx, sy = float(new_width) / width, float(new_height) / height
        self.fx, self.fy, self.cx, self.cy = fx * sx, fy * sy, cx * sx, cy * sy

    def _crop(self, a):
        h, w = a.shape[:2]
        return a[self.crop_y:h - self.crop_y, self.crop_x:w - self.crop_x]

    def apply_depth(self, depth):
        return cv2.resize(self._crop(depth), (self.new_width, self.new_height), interpolation=cv2.INTER_NEAREST)

    def apply_rgb(self, image, scale_rgb, mean_rgb, std_rgb, normalize_colors=True):
        out = cv2.resize(self._crop(image), (self.new_width, self.new_height), interpolation=cv2.INTER_LINEAR)
        if normalize_colors:
            out = out / scale_rgb
            for c in range(3):
                out[:, :, c] = (out[:, :, c] - mean_rgb[c]) / std_rgb[c]
        return out

    def apply_rgb_cuda(self, image, scale_rgb, mean_rgb, std_rgb, normalize_colors=True, out=None):
        """Device version of apply_rgb + the script's transpose / float / upload (fusionnet/run-testing.py:122-127).
        image: CUDA tensor (H,W,3) -- uint8 BGR (cv2.imread / load_image_u8) or float32 RGB (load_image).  Returns the
        (1,3,new_height,new_width) fp32 CUDA tensor the modules take.  Work is enqueued on the current stream."""
        from . import _ops
        return _ops.preprocess_rgb(image, self.crop_x, self.crop_y, self.new_height, self.new_width, scale_rgb, mean_rgb, std_rgb,
                                   normalize=normalize_colors, out=out)

    def get_updated_intrinsics(self):
        return np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]])
