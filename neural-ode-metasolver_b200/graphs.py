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

`HostFedLoop` is the host side of a loop that feeds such a step from pinned host batches: the step's scalar result
(loss / metric) is copied to a pinned slot asynchronously and handed back `lag` calls later, so the host never waits for
the step it has just enqueued and the GPU never idles on a graph launch (~1.5 ms for the 2 600-node premetanode10 step).
"""
import collections

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


class HostFedLoop:
    """step_fn(*device_inputs) -> 0-d device tensor.  `loop(x_host, y_host)`
      * copies the pinned host tensors into one of `lag + 1` device staging sets on a COPY stream (the copy of step k+1
        overlaps the kernels of step k; the compute stream waits on the copy's event),
      * enqueues the step on the current stream (GraphedStep.__call__ moves the staged batch into the graph's static
        inputs with a device-to-device copy and replays),
      * enqueues the device->host copy of the step's result into a pinned slot,
    and returns the result of the call made `lag` calls earlier as a Python float (None for the first `lag` calls);
    `drain()` returns the outstanding ones in order.  Every result is read on the host, none before its step has
    finished, and a staging set is not overwritten before the step that read it has finished."""

    def __init__(self, step_fn, lag=1):
        self.step_fn = step_fn
        self.lag = int(lag)
        n = self.lag + 1
        self._slots = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(n)]
        self._done = [torch.cuda.Event() for _ in range(n)]       # step finished, result in the pinned slot
        self._ready = [torch.cuda.Event() for _ in range(n)]      # staging set filled
        self._stage = [None] * n
        self._copy_stream = torch.cuda.Stream()
        self._pending = collections.deque()
        self._n = 0

    def _pop(self):
        j = self._pending.popleft()
        self._done[j].synchronize()
        return float(self._slots[j])

    def __call__(self, *host_inputs):
        i = self._n % len(self._slots)
        first_use = self._stage[i] is None
        if first_use:
            self._stage[i] = [torch.empty_strided(h.shape, h.stride(), dtype=h.dtype, device="cuda") for h in host_inputs]
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            if first_use:
                self._copy_stream.wait_stream(cur)                # the allocation above belongs to the compute stream
            else:
                self._copy_stream.wait_event(self._done[i])       # the step that read this staging set has finished
            for d, h in zip(self._stage[i], host_inputs):
                d.copy_(h, non_blocking=True)
            self._ready[i].record()
        cur.wait_event(self._ready[i])
        out = self.step_fn(*self._stage[i])
        self._n += 1
        self._slots[i].copy_(out.detach().reshape(()).float(), non_blocking=True)
        self._done[i].record()
        self._pending.append(i)
        return self._pop() if len(self._pending) > self.lag else None

    def drain(self):
        return [self._pop() for _ in range(len(self._pending))]
