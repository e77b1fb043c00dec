"""TEST INFRASTRUCTURE (imported only by tests/): CPU restatement of the quantisation the reference applies when it
writes its results -- `torchvision.utils.save_image(imgs, path, normalize=True, range=(-1, 1))` at
attention/run_attention.py:1470 and 1535, mapper/scripts/inference.py:74.

torchvision is a third-party dependency of the reference (not vendored under /root/reference; 0.26.0 is installed in
this image, where the keyword is spelled `value_range`).  Its published algorithm, torchvision/utils.py:
    make_grid(..., normalize=True, value_range=(low, high)):   norm_ip:  img.clamp_(min=low, max=high)
                                                                         img.sub_(low).div_(max(high - low, 1e-5))
    save_image:   ndarr = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8).numpy()
i.e. five fp32 operations in that order and a truncating conversion.  `Generator.set_image_output(torch.uint8)` performs
them in the last layer's epilogue (csrc/modconv_tc2.cu: quant_u8).  Pinned against torchvision itself by
tests/test_next_rows_host.py::test_save_image_oracle_equals_torchvision.
"""
import numpy as np


def quantise_ref(img, low=-1.0, high=1.0):
    """fp32 array (any shape) -> uint8 of the same shape, torchvision's operation order in fp32."""
    t = np.asarray(img, dtype=np.float32).copy()
    np.clip(t, np.float32(low), np.float32(high), out=t)                    # norm_ip: clamp_
    t = (t - np.float32(low)) / np.float32(max(high - low, 1e-5))          # sub_(low).div_(...)
    t = t * np.float32(255) + np.float32(0.5)                              # save_image: mul(255).add_(0.5)
    np.clip(t, np.float32(0), np.float32(255), out=t)                      # clamp_(0, 255)
    return t.astype(np.uint8)                                              # .to(torch.uint8): truncation
