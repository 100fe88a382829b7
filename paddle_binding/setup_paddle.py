"""Build the Paddle custom-op library on a machine WITH PaddlePaddle >= 2.1 (not possible in this image: no network, no wheel).

    LWS_ROOT=/path/to/this/repo python paddle_binding/setup_paddle.py install

Produces the importable module `lws_paddle_ops`; liblws_b200.so (make -C lwsnet_b200/csrc) must be on the loader path."""
import os

from paddle.utils.cpp_extension import CUDAExtension, setup  # noqa: E402  (needs Paddle)

root = os.environ.get("LWS_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
setup(name="lws_paddle_ops",
      ext_modules=CUDAExtension(sources=[os.path.join(root, "paddle_binding", "lws_paddle_ops.cc")],
                                include_dirs=[os.path.join(root, "include")],
                                library_dirs=[os.path.join(root, "lwsnet_b200", "lib")],
                                libraries=["lws_b200"]))
