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


class BaseGanExperiment(object):
    @property
    def generator(self):
        raise NotImplementedError()

    @property
    def discriminator(self):
        raise NotImplementedError()

    @property
    def generator_trainer(self):
        raise NotImplementedError()

    @property
    def discriminator_trainer(self):
        raise NotImplementedError()

    @property
    def feature_spec(self):
        raise NotImplementedError()

    def from_audio(self, samples, sr):
        raise NotImplementedError()

    def audio_representation(self, data, sr):
        raise NotImplementedError()

    def preprocess_batch(self, batch):
        return batch

    def to(self, device):
        self.generator.to(device)
        self.discriminator.to(device)
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
    def _gen_name(cls, prefix=''):
        return f'{cls.CHECKPOINT_DIR}/{prefix}{cls._name()}_gen.dat'

    @classmethod
    def _disc_name(cls, prefix=''):
        return f'{cls.CHECKPOINT_DIR}/{prefix}{cls._name()}_disc.dat'

    @classmethod
    def load_generator_weights(cls, generator, prefix=''):
        generator.load_state_dict(torch.load(cls._gen_name(prefix), map_location='cpu'))
        return generator

    @staticmethod
    def _host_state(module):
        # parameters are views of one flat buffer: clone, or every file would carry all of it
        return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}

    def checkpoint(self, prefix=''):
        os.makedirs(self.CHECKPOINT_DIR, exist_ok=True)
        torch.save(self._host_state(self.generator), self._gen_name(prefix))
        torch.save(self._host_state(self.discriminator), self._disc_name(prefix))

    def resume(self, prefix=''):
        self.generator.load_state_dict(torch.load(self._gen_name(prefix), map_location='cpu'))
        self.discriminator.load_state_dict(torch.load(self._disc_name(prefix), map_location='cpu'))
        for optim in (self.__g_optim, self.__d_optim):
            if optim is not None:
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
