"""Adam with the reference's configuration (featuresynth/experiment/experiment.py:111-117:
`Adam(parameters, lr=1e-4, betas=(0.5, 0.9))`, eps 1e-8, no weight decay / amsgrad) as ONE
kernel launch per step over a flat parameter buffer.

Construction re-points every parameter (and its .grad) at a slice of one contiguous fp32
buffer -- the layout a data-parallel step wants anyway: the gradient all-reduce of the network
being stepped is a single NCCL call over NVLink on the flat gradient buffer
(featuresynth trains single-device; SURVEY section 8(e) defines the data-parallel contract:
sum over ranks, then 1/world_size because the losses are batch means).
"""
import torch
import torch.distributed as dist

from .. import grad_ops


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None,
                 distributed=True):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise ValueError("parameters must live on a CUDA device (no CPU path)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.process_group = process_group
        # False: never reduce / broadcast (a single-process copy inside a data-parallel job)
        self.distributed = bool(distributed)
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # device-side step counter
        self.found_inf = torch.zeros(1, dtype=torch.int32, device=dev)  # set by unscale_()
        self._skip = None
        n = sum(p.numel() for p in self.params)
        # 16-byte aligned slices so the packed-weight kernels' vector loads stay aligned
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.offs = offs
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        self.numel = n
        self.sync_parameters()

    def sync_parameters(self):
        """data-parallel replicas must start from the same weights (`weights_init` is random and
        runs on every rank; ranks may resume from different files): broadcast rank 0's flat
        parameter and moment buffers.  Called at construction and after load_state_dict."""
        if self.world_size() > 1:
            for t in (self.flat, self.exp_avg, self.exp_avg_sq):
                dist.broadcast(t, src=dist.get_global_rank(self.process_group, 0)
                               if self.process_group is not None else 0, group=self.process_group)
            dist.broadcast(self.step_dev, src=dist.get_global_rank(self.process_group, 0)
                           if self.process_group is not None else 0, group=self.process_group)
            self.mark_updated()

    def _grad_view(self, i):
        p, o = self.params[i], self.offs[i]
        return self.flat_grad[o:o + p.numel()].view_as(p)

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        for i, p in enumerate(self.params):
            # someone may have dropped the views (module.zero_grad(), p.grad = None): the buffer
            # was just zeroed, so re-point without copying anything back
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * self.offs[i]:
                p.grad = self._grad_view(i)

    def _check_views(self):
        """step()/unscale_() read the flat buffers: a gradient that autograd allocated afresh
        (the view was dropped after zero_grad) is copied in; a parameter that no longer lives in
        the flat buffer (`.to()`, `.half()` after construction) cannot be stepped -> raise."""
        base_g, base_p = self.flat_grad.data_ptr(), self.flat.data_ptr()
        for i, p in enumerate(self.params):
            if p.data_ptr() != base_p + 4 * self.offs[i]:
                raise RuntimeError(
                    "parameter %d was moved out of the optimizer's flat buffer (module.to()/"
                    ".half() after constructing Adam); rebuild the optimizer" % i)
            if p.grad is None:
                p.grad = self._grad_view(i)          # no gradient this step: stays zero
            elif p.grad.data_ptr() != base_g + 4 * self.offs[i]:
                view = self._grad_view(i)
                view.copy_(p.grad)
                p.grad = view

    def world_size(self):
        if self.distributed and dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def all_reduce_grads(self):
        """sum the flat gradient over the data-parallel ranks (one collective)"""
        if self.world_size() > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.process_group)

    def unscale_(self, inv_scale_dev):
        """loss-scaled backward: reduce over ranks, multiply the flat gradient by the device
        scalar `inv_scale_dev` in place and record whether any element is non-finite; the next
        step() is then skipped on the device (no host round trip)"""
        self._check_views()
        if self.world_size() > 1:
            self.all_reduce_grads()
        grad_ops.grad_unscale_check(self.flat_grad, inv_scale_dev, self.found_inf)
        self._skip = self.found_inf
        self._reduced = True

    def step(self):
        world = self.world_size()
        if not getattr(self, "_reduced", False):
            self._check_views()
            if world > 1:
                self.all_reduce_grads()
        self._reduced = False
        self.step_count += 1
        grad_ops.adam_step_dev(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr,
                               self.betas[0], self.betas[1], self.eps, self.step_dev, 1.0 / world,
                               skip_flag=self._skip)
        self._skip = None
        self.mark_updated()

    def mark_updated(self):
        """the kernel wrote the parameters behind autograd's back: bump the version counters so
        that the packed 16-bit weight images are rebuilt (also called after a graph replay)"""
        for p in self.params:
            torch.autograd.graph.increment_version(p)

    def state_dict(self):
        return {"step": int(self.step_dev.item()), "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "lr": self.lr, "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.step_dev.fill_(self.step_count)
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.sync_parameters()
