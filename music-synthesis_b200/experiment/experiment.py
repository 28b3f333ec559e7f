"""Drop-in mirror of the experiment layer that drives the hot path
(featuresynth/experiment/experiment.py:9-221): `BaseGanExperiment` / `Experiment` wire a
generator and a discriminator to the two trainers, Adam(lr, betas=(0.5, 0.9)), the initialiser
contract, the alternating `training_steps` (discriminator first), checkpoint / resume in the
reference's file format (plain `state_dict()` pickles under trained_models/), the feature /
inference specs and batch pre-processing.

Differences, all at the edges of the path:
  * the optimisers are the flat-buffer Adam of train/optim.py and need their parameters in HBM,
    so they (and the trainers) are built when the experiment reaches a CUDA device -- `to(device)`
    or the first use of a trainer -- not in `__init__`;
  * `batch_stream(store, batch_size)` draws from a `DeviceAudioStore` (data/datastore.py) instead
    of walking a directory of audio files through LMDB-cached feature functions; `feature_funcs`
    is kept in the signature (and still required, as in the reference) but is informational;
  * reporting (HTML / S3) is out of scope.
"""
import os
from itertools import cycle

import torch

from ..data import batch_stream as device_batch_stream
from ..loss import hinge_discriminator_loss, hinge_generator_loss
from ..train import DiscriminatorTrainer, GeneratorTrainer
from ..train.optim import Adam
from .init import weights_init


def _required(name):
    def getter(self):
        raise NotImplementedError(name)
    return property(getter, doc="experiments must provide `%s`" % name)


class BaseGanExperiment(object):
    """What the training loop and the reporting code of the reference ask of an experiment."""
    generator = _required("generator")
    discriminator = _required("discriminator")
    generator_trainer = _required("generator_trainer")
    discriminator_trainer = _required("discriminator_trainer")
    feature_spec = _required("feature_spec")

    def from_audio(self, samples, sr):
        raise NotImplementedError("from_audio")

    def audio_representation(self, data, sr):
        raise NotImplementedError("audio_representation")

    def preprocess_batch(self, batch):
        return batch

    def to(self, device):
        for net in (self.generator, self.discriminator):
            net.to(device)
        return self


class Experiment(BaseGanExperiment):
    CHECKPOINT_DIR = "trained_models"

    def __init__(self, generator, discriminator, learning_rate, feature_size, audio_repr_class,
                 generator_loss, discriminator_loss, g_init=weights_init, d_init=weights_init,
                 feature_funcs=None, total_samples=16384, feature_channels=256,
                 inference_sequence_factor=4, samplerate=11025,
                 sub_disc_loss=hinge_discriminator_loss, sub_gen_loss=hinge_generator_loss,
                 cuda_graph=False, process_group=None):
        super().__init__()
        self.sub_gen_loss = sub_gen_loss
        self.sub_disc_loss = sub_disc_loss
        self.inference_sequence_factor = inference_sequence_factor
        if feature_funcs is None:
            raise ValueError('You must provide feature funcs')
        self.discriminator_init = d_init
        self.generator_init = g_init
        for net in (generator, discriminator):
            if hasattr(net, 'initialize_weights'):
                raise ValueError('initialize_weights() is deprecated: pass g_init / d_init')
        self.__g = generator
        self.__g.apply(g_init)
        self.__d = discriminator
        self.__d.apply(d_init)
        self.learning_rate = learning_rate
        self.generator_loss = generator_loss
        self.discriminator_loss = discriminator_loss
        self.cuda_graph = cuda_graph
        self.process_group = process_group
        self.__g_optim = self.__d_optim = None
        self.__g_trainer = self.__d_trainer = None
        self.__feature_size = feature_size
        self.__audio_repr_class = audio_repr_class
        self.__anchor_feature = 'spectrogram'
        self.__feature_funcs = feature_funcs
        self.training_steps = cycle([
            lambda *batch: self.discriminator_trainer(*batch),
            lambda *batch: self.generator_trainer(*batch)])
        self.samplerate = samplerate
        self.total_samples = total_samples
        self.feature_channels = feature_channels

    # -- trainers: built once the parameters are in HBM ------------------------------------
    def _build_trainers(self):
        if self.__g_trainer is not None:
            return
        first = next(self.__g.parameters())
        if first.device.type != 'cuda':
            self.to('cuda')
        betas = (0.5, 0.9)                                   # experiment.py:111-117
        self.__g_optim = Adam(self.__g.parameters(), lr=self.learning_rate, betas=betas,
                              process_group=self.process_group)
        self.__d_optim = Adam(self.__d.parameters(), lr=self.learning_rate, betas=betas,
                              process_group=self.process_group)
        self.__g_trainer = GeneratorTrainer(
            self.__g, self.__g_optim, self.__d, self.__d_optim, self.generator_loss,
            self.sub_gen_loss, cuda_graph=self.cuda_graph)
        self.__d_trainer = DiscriminatorTrainer(
            self.__g, self.__g_optim, self.__d, self.__d_optim, self.discriminator_loss,
            self.sub_disc_loss, cuda_graph=self.cuda_graph)

    def to(self, device):
        if self.__g_trainer is not None:
            raise RuntimeError('the optimisers own the parameter storage: move the experiment '
                               'before the first training step')
        return super().to(device)

    # -- checkpoints: the reference's files -------------------------------------------------
    @classmethod
    def _name(cls):
        return cls.__name__.lower().replace('experiment', '')

    @classmethod
    def _path(cls, role, prefix):
        """trained_models/{prefix}{name}_{gen|disc}.dat -- the reference's file names"""
        return '%s/%s%s_%s.dat' % (cls.CHECKPOINT_DIR, prefix, cls._name(), role)

    @classmethod
    def _gen_name(cls, prefix=''):
        return cls._path('gen', prefix)

    @classmethod
    def _disc_name(cls, prefix=''):
        return cls._path('disc', prefix)

    def _networks(self):
        return (('gen', self.generator), ('disc', self.discriminator))

    @classmethod
    def load_generator_weights(cls, generator, prefix=''):
        generator.load_state_dict(torch.load(cls._path('gen', prefix), map_location='cpu'))
        return generator

    def checkpoint(self, prefix=''):
        """plain state-dict pickles, as the reference writes them.  Parameters are views of one
        flat optimiser buffer: they are cloned, or every file would carry all of it."""
        os.makedirs(self.CHECKPOINT_DIR, exist_ok=True)
        for role, net in self._networks():
            host = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
            torch.save(host, self._path(role, prefix))

    def resume(self, prefix=''):
        for role, net in self._networks():
            net.load_state_dict(torch.load(self._path(role, prefix), map_location='cpu'))
        for optim in (self.__g_optim, self.__d_optim):
            if optim is not None:        # the kernels cache packed weights by tensor version
                optim.mark_updated()

    # -- the reference's accessors -----------------------------------------------------------
    @property
    def generator(self):
        return self.__g

    @property
    def discriminator(self):
        return self.__d

    @property
    def generator_trainer(self):
        self._build_trainers()
        return self.__g_trainer.train

    @property
    def discriminator_trainer(self):
        self._build_trainers()
        return self.__d_trainer.train

    @property
    def optimizers(self):
        self._build_trainers()
        return self.__g_optim, self.__d_optim

    def from_audio(self, samples, sr):
        return self.__audio_repr_class.from_audio(samples, sr)

    def audio_representation(self, data, sr):
        return self.__audio_repr_class(data, sr)

    def preprocess_batch(self, batch):
        samples, features = batch
        r = self.from_audio(samples, self.samplerate)
        return r.data, features

    def batch_stream(self, store, batch_size, feature_spec=None, seed=None, rank=0):
        return device_batch_stream(store, batch_size, feature_spec or self.feature_spec,
                                   self.__anchor_feature, seed=seed, rank=rank)

    @property
    def feature_funcs(self):
        return self.__feature_funcs

    @property
    def feature_spec(self):
        return {
            'audio': (self.total_samples, 1),
            'spectrogram': (self.__feature_size, self.feature_channels)
        }

    @property
    def inference_spec(self):
        return {k: (size * self.inference_sequence_factor, channels)
                for k, (size, channels) in self.feature_spec.items()}
