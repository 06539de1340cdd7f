"""CUDA-graph replay of ``Generator.forward`` for serving (configs[2]: latents + cameras -> 256^2 images).

One eager pass of the full generator is ~230 kernel launches, a third of them sub-10 us torch kernels of the mapping network and
the style heads: at B = 64 the GPU idles ~1 ms of a 13.7 ms pass waiting for launches.  Everything on the path is capturable -- the
C-ABI kernels launch on the caller's stream, tensor maps travel as kernel parameters, nothing synchronises or allocates outside
torch's caching allocator -- so the pass is captured once per input shape and replayed.
"""
import torch

__all__ = ["GraphedGenerator"]


class GraphedGenerator:
    """``gg = GraphedGenerator(generator, [z], cam_poses, focals, near, far, **forward_kwargs)``; ``gg([z], cam_poses, focals, near, far)``
    returns what ``generator(...)`` returns, from static output buffers that the next call overwrites.  Inference only (the generator
    must not require gradients: ``ema=True`` / ``eval()`` under ``torch.no_grad``); the keyword arguments are frozen at capture.
    With ``randomize_noise=True`` (the default) every replay draws fresh noise, as the eager call does."""

    def __init__(self, generator, styles, cam_poses, focals, near, far, warmup=2, **forward_kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedGenerator needs a CUDA device")
        self.generator = generator
        self.kwargs = dict(forward_kwargs)
        self._in = [[s.detach().clone() for s in styles]] + [t.detach().clone() if torch.is_tensor(t) else t for t in (cam_poses, focals, near, far)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):                 # builds the library, sizes the workspaces, fills every cache
                generator(*self._in, **self.kwargs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._out = generator(*self._in, **self.kwargs)

    def __call__(self, styles, cam_poses, focals, near, far):
        for dst, src in zip(self._in[0], styles):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(self._in[1:], (cam_poses, focals, near, far)):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
            elif dst != src:
                raise ValueError("GraphedGenerator: non-tensor argument differs from the captured one (%r vs %r)" % (src, dst))
        self.graph.replay()
        return self._out
