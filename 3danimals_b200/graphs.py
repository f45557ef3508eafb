"""CUDA-graph capture for the launch-bound loops of the path that keep their shapes: rotation renders and the texture
finetune of visualize_results.py (BASELINE configs[4]; reference visualization/visualize_results.py:353-396 renders 75
views x 3 mode sets of ONE fixed mesh, and its texture finetune runs N identical fwd+bwd iterations on it).

In training the extracted topology changes every step and the reference API returns exactly-sized tensors (one 16-byte
size readback per extraction), so whole-step graphs do not apply there.  With the geometry fixed, a frame is ~25 kernel
launches of libb2a.so plus a handful of PyTorch ones and the host needs ~0.8 ms to enqueue what the B200 executes in
~0.35 ms; captured once, the same work replays with a single launch.

`CapturedStep(fn, example_inputs)` captures `fn(*static_inputs)` (any mix of libb2a ops and PyTorch ops, forward only or
forward + torch.autograd.grad) into a torch.cuda.CUDAGraph; calling it copies new inputs into the static buffers, replays
and returns the static outputs (valid until the next call).  Requirements are CUDA-graph's own: no host synchronisation
inside `fn` (so the sparse covered-pixel field evaluation, which reads a row count, must be off: `sparse_fields=False`
or a `dense_only` field) and fixed shapes.
"""
import torch


class CapturedStep:
    def __init__(self, fn, example_inputs, warmup=3):
        if not all(torch.is_tensor(t) and t.is_cuda for t in example_inputs):
            raise RuntimeError("CapturedStep needs CUDA tensors (the B200 hot path has no CPU fallback)")
        self.fn = fn
        self.static_inputs = [t.detach().clone().requires_grad_(t.requires_grad) for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # warm-up off the capture: lazy workspaces, cuBLAS handles, autograd buffers
            for _ in range(max(int(warmup), 1)):
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_outputs = fn(*self.static_inputs)
        self.replays = 0

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_inputs):
            raise TypeError("expected %d inputs" % len(self.static_inputs))
        with torch.no_grad():
            for dst, src in zip(self.static_inputs, inputs):
                if dst.shape != src.shape:
                    raise RuntimeError("CapturedStep: input shape %s differs from the captured %s" % (tuple(src.shape), tuple(dst.shape)))
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src)
        self.graph.replay()
        self.replays += 1
        return self.static_outputs


def captured_render(mesh, prior_mesh, material, lgt, resolution, example_cameras, spp=1, render_modes=("shaded",), background=None,
                    dino_net=None, feat=None, class_vector=None, two_sided_shading=True):
    """Forward-only render of a FIXED mesh under changing cameras (rotation renders): returns f(mvp, w2c, campos) -> the list
    `render_mesh` returns, replayed from one CUDA graph."""
    from .render import render as render_mod
    mesh.edge_adjacency()        # topology tables are built outside the capture

    def frame(mvp, w2c, campos):
        with torch.no_grad():
            return render_mod.render_mesh(None, mesh, mvp, w2c, campos, material, lgt, resolution, spp=spp, num_layers=1, msaa=True,
                                          background=background, bsdf="diffuse", feat=feat, render_modes=list(render_modes),
                                          prior_mesh=prior_mesh, dino_net=dino_net, class_vector=class_vector,
                                          two_sided_shading=two_sided_shading, sparse_fields=False)

    return CapturedStep(frame, list(example_cameras))
