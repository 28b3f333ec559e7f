"""Drop-in mirrors of featuresynth/train/train.py:8-74 (GeneratorTrainer / DiscriminatorTrainer):
same constructor arguments, `train(samples, features)` contract and return dictionaries.

One optimiser step each, same arithmetic as the reference; two pieces of work the reference does
and then throws away are skipped:
  * DiscriminatorTrainer: the reference back-propagates through the generator and discards
    those gradients (zero_grad at the next step, App. E.4) -- here the fake batch comes from the
    fused inference kernels under no_grad, and the fake and real batches go through the
    discriminator in ONE pass (concatenated along the batch axis; same per-clip arithmetic);
  * GeneratorTrainer: the discriminator's weight gradients are not formed (its parameters are
    frozen for the call) and the real batch runs without a tape.
`exact_reference_grads=True` restores the reference's behaviour (all .grad fields populated).
Data-parallel: when torch.distributed is initialised, the gradients of the network being
stepped are summed over ranks before the step (optimisers from .optim do it in one NCCL call).

`cuda_graph=True` (needs the optimisers from .optim): after two eager calls per input shape
zero_grad + forward + backward (~320 kernel launches per step) are captured into ONE CUDA graph
and replayed; the gradient all-reduce (NCCL, kept outside the capture so that no collective
depends on capture-mode restrictions), the fused Adam kernel and the loss read-back follow
eagerly.
"""
import contextlib

import torch
import torch.distributed as dist

from .. import grad_ops
from ..loss.loss import hinge_discriminator_loss, hinge_generator_loss
from ..util.modules import zero_grad


class LossScaler:
    """Dynamic loss scaling, active only with MSB_GRAD_FMT=f16 (the default bf16 backward needs
    none; see grad_ops.GRAD_FMT).  Same policy as torch.amp.GradScaler:
    the backward pass is seeded with `scale` instead of 1, the flat gradient is multiplied by
    1/scale before the step; a non-finite gradient skips the step (on the device) and halves the
    scale, `growth_interval` clean steps double it.  The scale lives in device memory so that a
    captured CUDA graph reads the current value."""

    def __init__(self, device, init_scale=2.0 ** 16, growth_interval=500, max_scale=2.0 ** 30):
        self.enabled = grad_ops.NEEDS_LOSS_SCALE
        self.scale = float(init_scale) if self.enabled else 1.0
        self.growth_interval = growth_interval
        self.max_scale = max_scale
        self.clean = 0
        self.skipped = 0
        self.scale_dev = torch.full((), self.scale, dtype=torch.float32, device=device)
        self.inv_scale_dev = torch.full((1,), 1.0 / self.scale, dtype=torch.float32, device=device)

    def _set(self, scale):
        self.scale = scale
        self.scale_dev.fill_(scale)
        self.inv_scale_dev.fill_(1.0 / scale)

    def update(self, found_inf):
        if not self.enabled:
            return
        if found_inf:
            self.skipped += 1
            self.clean = 0
            self._set(max(self.scale * 0.5, 1.0))
        else:
            self.clean += 1
            if self.clean >= self.growth_interval and self.scale < self.max_scale:
                self.clean = 0
                self._set(self.scale * 2.0)


@contextlib.contextmanager
def _frozen(module):
    flags = [p.requires_grad for p in module.parameters()]
    for p in module.parameters():
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p, f in zip(module.parameters(), flags):
            p.requires_grad_(f)


def _sync_grads(optim, module):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    if hasattr(optim, "all_reduce_grads"):
        return                                   # .optim.Adam reduces inside step()
    if not getattr(optim, "distributed", True):
        return
    world = dist.get_world_size()
    for p in module.parameters():
        if p.grad is not None:
            dist.all_reduce(p.grad)
            p.grad.div_(world)


class _Graphed:
    """captured step for one input shape: static input buffers, the graph, static outputs"""

    def __init__(self):
        self.calls = 0
        self.graph = None
        self.samples = self.features = None
        self.outputs = None
        self.launches = 0           # kernel launches recorded in the graph


class _GraphMixin:
    def _init_graph(self, cuda_graph):
        self.cuda_graph = cuda_graph
        self._graphs = {}
        #: library kernel launches inside the most recently captured step graph
        self.graph_launches = 0

    def _all_params(self):
        return list(self.generator.parameters()) + list(self.discriminator.parameters())

    #: initial loss scale of this trainer's backward (None: LossScaler default)
    init_loss_scale = None

    def _scaler(self, device):
        if getattr(self, "scaler", None) is None:
            kw = {} if self.init_loss_scale is None else {"init_scale": self.init_loss_scale}
            self.scaler = LossScaler(device, **kw)
        return self.scaler

    def _finish(self, optim, module):
        """after backward: data-parallel reduction, un-scaling + overflow check, optimiser step,
        loss-scale update"""
        sc = self.scaler
        if hasattr(optim, "unscale_"):                    # .optim.Adam: all on the device
            if sc.enabled:
                optim.unscale_(sc.inv_scale_dev)
            optim.step()
            if sc.enabled:
                sc.update(bool(optim.found_inf.item()))
            return
        _sync_grads(optim, module)                        # any torch optimiser
        bad = False
        if sc.enabled:
            for p in module.parameters():
                if p.grad is not None:
                    p.grad.mul_(1.0 / sc.scale)
                    bad = bad or not bool(torch.isfinite(p.grad).all())
        if not bad:
            optim.step()
        sc.update(bad)

    def _run(self, samples, features):
        """eager or graphed zero_grad + forward + backward -> tuple of output tensors"""
        if not self.cuda_graph:
            return self._fwd_bwd(samples, features)
        for o in (self.g_optim, self.d_optim):
            if not hasattr(o, "flat_grad"):
                raise ValueError("cuda_graph=True needs the flat-buffer optimisers of "
                                 "music_synthesis_b200.train.optim (gradients must keep their "
                                 "addresses across replays)")
        multi = isinstance(samples, dict)            # MultiScale audio: {band size: tensor}
        key = (tuple((k, tuple(v.shape)) for k, v in samples.items()) if multi
               else tuple(samples.shape), tuple(features.shape))
        st = self._graphs.setdefault(key, _Graphed())
        st.calls += 1
        if st.calls <= 2:                       # warm-up: lazy initialisation, allocator
            return self._fwd_bwd(samples, features)
        if st.graph is None:
            st.samples = {k: v.clone() for k, v in samples.items()} if multi else samples.clone()
            st.features = features.clone()
            for p in self._all_params():        # every packed-weight image is rebuilt IN the graph
                torch.autograd.graph.increment_version(p)
            torch.cuda.synchronize()
            st.graph = torch.cuda.CUDAGraph()
            from .. import _lib
            n0 = _lib.launch_count()
            with torch.cuda.graph(st.graph):
                st.outputs = self._fwd_bwd(st.samples, st.features)
            st.launches = _lib.launch_count() - n0
            self.graph_launches = st.launches
        if multi:
            for k, v in samples.items():
                st.samples[k].copy_(v)
        else:
            st.samples.copy_(samples)
        st.features.copy_(features)
        st.graph.replay()
        return st.outputs


class GeneratorTrainer(_GraphMixin):
    # the generator's gradient passes through the discriminator and batch means over whole
    # feature maps: activation gradients of 1e-6 .. 1e-9
    init_loss_scale = 2.0 ** 22

    def __init__(self, generator, g_optim, discriminator, d_optim, loss,
                 sub_loss=hinge_generator_loss, exact_reference_grads=False, cuda_graph=False):
        super().__init__()
        self._init_graph(cuda_graph)
        self.sub_loss = sub_loss
        self.loss = loss
        self.d_optim = d_optim
        self.discriminator = discriminator
        self.g_optim = g_optim
        self.generator = generator
        self.exact_reference_grads = exact_reference_grads

    def _fwd_bwd(self, samples, features):
        zero_grad(self.g_optim, self.d_optim)
        fake = self.generator(features)
        if self.exact_reference_grads:
            f_features, f_score = self.discriminator(fake, features)
            r_features, r_score = self.discriminator(samples, features)
        else:
            with _frozen(self.discriminator):
                f_features, f_score = self.discriminator(fake, features)
                with torch.no_grad():
                    r_features, r_score = self.discriminator(samples, features)
        loss = self.loss(r_features, f_features, r_score, f_score, gan_loss=self.sub_loss)
        loss.backward(self._scaler(loss.device).scale_dev)
        fake = {k: v.detach() for k, v in fake.items()} if isinstance(fake, dict) else fake.detach()
        return loss.detach(), fake

    def train(self, samples, features):
        loss, fake = self._run(samples, features)
        self._finish(self.g_optim, self.generator)
        try:
            fake = fake.data.cpu().numpy()
        except AttributeError:
            fake = {k: v.data.cpu().numpy() for k, v in fake.items()}
        return {'g_loss': loss.item(), 'fake': fake}


class DiscriminatorTrainer(_GraphMixin):
    def __init__(self, generator, g_optim, discriminator, d_optim, loss,
                 sub_loss=hinge_discriminator_loss, exact_reference_grads=False, cuda_graph=False):
        super().__init__()
        self._init_graph(cuda_graph)
        self.sub_loss = sub_loss
        self.loss = loss
        self.d_optim = d_optim
        self.discriminator = discriminator
        self.g_optim = g_optim
        self.generator = generator
        self.exact_reference_grads = exact_reference_grads

    def _fwd_bwd(self, samples, features):
        zero_grad(self.g_optim, self.d_optim)
        if self.exact_reference_grads:
            fake = self.generator(features)
        else:
            from .. import ops
            with torch.no_grad(), ops.relaxed_forward():
                fake = self.generator(features)
        if self.exact_reference_grads:
            _, f_score = self.discriminator(fake, features)
            _, r_score = self.discriminator(samples, features)
        else:
            # one pass over [fake; real]: the discriminators have no cross-clip operation, so this
            # is the same arithmetic per clip with half the kernel launches
            B = features.shape[0]
            if isinstance(fake, dict):
                both = {k: torch.cat([fake[k], samples[k]], dim=0) for k in fake}
            else:
                both = torch.cat([fake, samples], dim=0)
            _, score = self.discriminator(both, torch.cat([features, features], dim=0))
            f_score = [j[:B] for j in score]
            r_score = [j[B:] for j in score]
        loss = self.loss(r_score, f_score, gan_loss=self.sub_loss)
        loss.backward(self._scaler(loss.device).scale_dev)
        return (loss.detach(),)

    def train(self, samples, features):
        (loss,) = self._run(samples, features)
        self._finish(self.d_optim, self.discriminator)
        return {'d_loss': loss.item()}


def training_loop(batch_stream, experiment, device, loggers):
    """featuresynth/train/train.py:77-106: pre-process each batch, run the experiment's next
    training step (discriminator and generator alternate), hand the result to the loggers.
    Batches may already be CUDA tensors (data/datastore.py) -- then nothing is copied."""
    from datetime import datetime, timezone

    def to_device(x):
        if isinstance(x, dict):
            return {k: to_device(v) for k, v in x.items()}
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(x)
        return x.to(device).float()

    start_time = datetime.now(timezone.utc)
    for i, batch in enumerate(batch_stream):
        preprocessed = experiment.preprocess_batch(batch)
        tensors = [to_device(x) for x in preprocessed]
        step = next(experiment.training_steps)
        step_result = step(*tensors)
        elapsed_time = datetime.now(timezone.utc) - start_time
        log_results = {}
        for logger in loggers:
            log_result = logger(experiment, preprocessed, step_result, i, elapsed_time)
            if log_result is not None:
                log_results.update(log_result)
        yield i, elapsed_time, log_results
