"""The multiscale experiments of featuresynth/experiment/multiscale.py that this path covers:
  FilterBankMultiscaleExperiment                  :17-63   (BASELINE config 5 wiring)
  MultiScaleNoDeRecomposeUnconditionedShortKernel :112-155
  MultiScaleNoDeRecompose                         :158-201
All train on MultiScale band dictionaries (recompose=False / decompose=False) with the
least-squares sub-losses."""
from ..audio.representation import MultiScale
from ..discriminator.multiscale import (FilterBankMultiScaleDiscriminator,
                                        MultiScaleMultiResDiscriminator)
from ..generator.multiscale import FilterBankMultiScaleGenerator, MultiScaleGenerator
from ..loss import (least_squares_disc_loss, least_squares_generator_loss, mel_gan_disc_loss,
                    mel_gan_gen_loss)
from .experiment import Experiment
from .init import weights_init

_FEATURE_FUNCS = {'audio': ('DeviceAudioStore.audio', (22050,)),
                  'spectrogram': ('Audio2Mel', (22050,))}


def _common(n_mels, feature_size, total_samples, samplerate):
    return dict(
        learning_rate=1e-4,
        feature_size=feature_size,
        audio_repr_class=MultiScale,
        generator_loss=mel_gan_gen_loss,
        sub_gen_loss=least_squares_generator_loss,
        discriminator_loss=mel_gan_disc_loss,
        sub_disc_loss=least_squares_disc_loss,
        g_init=weights_init,
        d_init=weights_init,
        feature_funcs=_FEATURE_FUNCS,
        total_samples=total_samples,
        feature_channels=n_mels,
        samplerate=samplerate,
        inference_sequence_factor=4)


class FilterBankMultiscaleExperiment(Experiment):
    AUDIO_REPR_CLASS = MultiScale
    SAMPLERATE = 22050
    N_MELS = 128
    feature_size = 32
    total_samples = 8192

    @classmethod
    def make_generator(cls):
        return FilterBankMultiScaleGenerator(
            cls.SAMPLERATE, cls.N_MELS, cls.feature_size, cls.total_samples, recompose=False)

    def __init__(self, **kw):
        super().__init__(
            generator=self.make_generator(),
            discriminator=FilterBankMultiScaleDiscriminator(
                self.total_samples, self.SAMPLERATE, decompose=False,
                conditioning_channels=self.N_MELS),
            **_common(self.N_MELS, self.feature_size, self.total_samples, self.SAMPLERATE), **kw)


class MultiScaleNoDeRecomposeUnconditionedShortKernel(Experiment):
    def __init__(self, **kw):
        n_mels, feature_size, total_samples = 128, 32, 8192
        super().__init__(
            generator=MultiScaleGenerator(
                n_mels, feature_size, total_samples, transposed_conv=True, recompose=False),
            discriminator=MultiScaleMultiResDiscriminator(
                total_samples, flatten_multiscale_features=False, channel_judgements=True,
                decompose=False, kernel_size=9),
            **_common(n_mels, feature_size, total_samples, 22050), **kw)


class MultiScaleNoDeRecompose(Experiment):
    def __init__(self, **kw):
        n_mels, feature_size, total_samples = 128, 32, 8192
        super().__init__(
            generator=MultiScaleGenerator(
                n_mels, feature_size, total_samples, transposed_conv=True, recompose=False),
            discriminator=MultiScaleMultiResDiscriminator(
                total_samples, flatten_multiscale_features=False, channel_judgements=True,
                conditioning_channels=n_mels, decompose=False),
            **_common(n_mels, feature_size, total_samples, 22050), **kw)
