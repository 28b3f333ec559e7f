"""The experiment wirings of the reference that this path covers, as data: which generator /
discriminator pair, which audio representation, which sub-losses.  Everything else is the
reference's common configuration (Adam lr 1e-4, weights_init on both networks, 8192-sample crops
with 32 log-mel frames of 128 bins at 22 050 Hz, inference on 4x longer sequences).

  MultiScaleMelGanExperiment                      featuresynth/experiment/melgan.py:11-46
  RealMelGanExperiment                            featuresynth/experiment/realmelgan.py:219-254
  FilterBankMultiscaleExperiment                  featuresynth/experiment/multiscale.py:17-63
  MultiScaleNoDeRecomposeUnconditionedShortKernel featuresynth/experiment/multiscale.py:112-155
  MultiScaleNoDeRecompose                         featuresynth/experiment/multiscale.py:158-201
  FilterBankExperiment                            featuresynth/experiment/filterbank.py:14-78
  ConditionalFilterBankExperiment                 featuresynth/experiment/filterbank.py:81-128
  AlternateFilterBankExperiment                   featuresynth/experiment/filterbank.py:131-206
"""
from ..audio.representation import MultiScale, RawAudio
from ..loss import (hinge_discriminator_loss, hinge_generator_loss, least_squares_disc_loss,
                    least_squares_generator_loss, mel_gan_disc_loss, mel_gan_gen_loss)
from .experiment import Experiment
from .init import weights_init

SAMPLERATE, N_FFT, HOP, N_MELS = 22050, 1024, 256, 128
TOTAL_SAMPLES = 8192
FEATURE_SIZE = TOTAL_SAMPLES // HOP

# what the reference computes through its LMDB-cached feature functions lives in the
# DeviceAudioStore here; the dict is informational (Experiment still insists on one)
FEATURE_FUNCS = {'audio': ('DeviceAudioStore.audio', (SAMPLERATE,)),
                 'spectrogram': ('Audio2Mel', (SAMPLERATE, N_FFT, HOP, N_MELS))}

HINGE = dict(sub_gen_loss=hinge_generator_loss, sub_disc_loss=hinge_discriminator_loss)
LEAST_SQUARES = dict(sub_gen_loss=least_squares_generator_loss,
                     sub_disc_loss=least_squares_disc_loss)


class _Wired(Experiment):
    """subclasses name their pair in `build()` and pick the representation / sub-losses"""
    representation = RawAudio
    sub_losses = HINGE
    generator_loss = staticmethod(mel_gan_gen_loss)

    def build(self):
        raise NotImplementedError()

    def __init__(self, **kw):
        generator, discriminator = self.build()
        super().__init__(
            generator, discriminator,
            learning_rate=1e-4,
            feature_size=FEATURE_SIZE,
            audio_repr_class=self.representation,
            generator_loss=type(self).generator_loss,
            discriminator_loss=mel_gan_disc_loss,
            g_init=weights_init, d_init=weights_init,
            feature_funcs=FEATURE_FUNCS,
            total_samples=TOTAL_SAMPLES,
            feature_channels=N_MELS,
            samplerate=SAMPLERATE,
            inference_sequence_factor=4,
            **self.sub_losses, **kw)


class MultiScaleMelGanExperiment(_Wired):
    def build(self):
        from ..discriminator import MelGanDiscriminator
        from ..generator.full import MelGanGenerator
        return MelGanGenerator(FEATURE_SIZE, N_MELS), MelGanDiscriminator()


class RealMelGanExperiment(_Wired):
    def build(self):
        from . import realmelgan as official
        return (official.Generator(N_MELS, FEATURE_SIZE, n_residual_layers=3),
                official.Discriminator(num_D=3, ndf=16, n_layers=4, downsampling_factor=4))

    @staticmethod
    def generator_loss(*a, **kw):
        # the pair uses ITS feature-matching weighting (realmelgan.py:205-217)
        from .realmelgan import mel_gan_gen_loss as official_gen_loss
        return official_gen_loss(*a, **kw)


class FilterBankMultiscaleExperiment(_Wired):
    representation = AUDIO_REPR_CLASS = MultiScale
    sub_losses = LEAST_SQUARES
    N_MELS = N_MELS
    feature_size = FEATURE_SIZE
    total_samples = TOTAL_SAMPLES

    @classmethod
    def make_generator(cls):
        from ..generator.multiscale import FilterBankMultiScaleGenerator
        return FilterBankMultiScaleGenerator(SAMPLERATE, N_MELS, FEATURE_SIZE, TOTAL_SAMPLES,
                                             recompose=False)

    def build(self):
        from ..discriminator.multiscale import FilterBankMultiScaleDiscriminator
        return self.make_generator(), FilterBankMultiScaleDiscriminator(
            TOTAL_SAMPLES, SAMPLERATE, decompose=False, conditioning_channels=N_MELS)


class _BandDictionaryPair(_Wired):
    """MultiScaleGenerator (transposed convs, bands out) + MultiScaleMultiResDiscriminator
    (bands in, per-channel judgements); subclasses choose the discriminator's options"""
    representation = MultiScale
    sub_losses = LEAST_SQUARES
    discriminator_options = {}

    def build(self):
        from ..discriminator.multiscale import MultiScaleMultiResDiscriminator
        from ..generator.multiscale import MultiScaleGenerator
        return (MultiScaleGenerator(N_MELS, FEATURE_SIZE, TOTAL_SAMPLES, transposed_conv=True,
                                    recompose=False),
                MultiScaleMultiResDiscriminator(TOTAL_SAMPLES, flatten_multiscale_features=False,
                                                channel_judgements=True, decompose=False,
                                                **self.discriminator_options))


class MultiScaleNoDeRecomposeUnconditionedShortKernel(_BandDictionaryPair):
    discriminator_options = dict(kernel_size=9)


class MultiScaleNoDeRecompose(_BandDictionaryPair):
    discriminator_options = dict(conditioning_channels=N_MELS)


class FilterBankExperiment(_Wired):
    """experiment/filterbank.py:14-78: FilterBankGenerator + (unconditioned)
    FilterBankDiscriminator over one fixed 511-tap, 128-band linear-scale Morlet bank
    (20 Hz .. nyquist - 20 Hz, scaling factor 0.9), least-squares sub-losses, raw audio"""
    sub_losses = LEAST_SQUARES
    N_MELS = N_MELS
    FEATURE_SIZE = FEATURE_SIZE
    TOTAL_SAMPLES = TOTAL_SAMPLES
    conditioning_channels = 0

    @classmethod
    def make_filter_bank(cls, samplerate=SAMPLERATE):
        from ..audio.filterbank import FilterBank, SampleRate, linear_center_frequencies
        sr = SampleRate(float(samplerate))
        return FilterBank(sr, 511, linear_center_frequencies(20, sr.nyquist - 20, 128),
                          scaling_factors=0.9, normalize_filters=True, a_weighting=False)

    @classmethod
    def make_generator(cls, filter_bank=None):
        from ..generator.filterbank import FilterBankGenerator
        return FilterBankGenerator(filter_bank or cls.make_filter_bank(), FEATURE_SIZE,
                                   TOTAL_SAMPLES, N_MELS)

    def build(self):
        from ..discriminator.filterbank import FilterBankDiscriminator
        return self.make_generator(), FilterBankDiscriminator(
            self.make_filter_bank(), TOTAL_SAMPLES,
            conditioning_channels=self.conditioning_channels)


class ConditionalFilterBankExperiment(FilterBankExperiment):
    """experiment/filterbank.py:81-128: the same pair with the discriminator conditioned on the
    128 log-mel channels"""
    conditioning_channels = N_MELS


class AlternateFilterBankExperiment(FilterBankExperiment):
    """experiment/filterbank.py:131-206: the MelGAN-shaped, weight-normed
    ResidualStackFilterBankGenerator (harmonic + noise heads) against the conditioned
    FilterBankDiscriminator, each over its own (identically configured) bank"""
    conditioning_channels = N_MELS

    @classmethod
    def make_generator(cls, filter_bank=None):
        from ..generator.filterbank import ResidualStackFilterBankGenerator
        return ResidualStackFilterBankGenerator(filter_bank or cls.make_filter_bank(), FEATURE_SIZE,
                                                TOTAL_SAMPLES, N_MELS, add_weight_norm=True)
