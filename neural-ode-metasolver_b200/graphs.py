"""CUDA-graph capture of a whole training / attack step.

One step of premetanode10 is ~255 kernel launches of 20-200 us each; issued eagerly from Python the GPU idles
~10 % of the step waiting for the host.  Every launch of this library is a plain stream-ordered kernel with
host-side scalars baked into its parameters (tableau coefficients, TMA descriptors), so a step whose solver
parameters are fixed can be captured ONCE and replayed with no host work:

    g = GraphedStep(lambda x, y: train_step(model, x, y), (x_example, y_example))
    loss = g(x, y)          # copies x, y into the static buffers, replays, returns the static loss tensor

Re-capture when u / v change (solver smoothing draws new values per batch: capture one graph per drawn value,
or run those steps eagerly).  Gradients produced inside the step live in the graph's memory pool and are
overwritten by the next replay.

Everything the HOST decides during the step is frozen into the graph, not only u / v:
  * the solver id drawn by the 'switch' regime (numpy RNG) and the coin flip of 'ensemble' with ensemble_prob < 1
    (CPU torch.bernoulli) -- a replay repeats the captured draw;
  * `FusedSGD.step()`'s lr, grad_scale and first-step flag (it refuses to be captured; keep the optimizer step, and a
    CyclicLR schedule, outside the graph);
  * the tableau-gradient backward (unfrozen solvers) copies 20 doubles to the host and cannot be captured at all.
"""
import torch


class GraphedStep:
    def __init__(self, fn, example_inputs, warmup=3):
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                     # warm-up off the default stream (allocator, autotune-free)
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
