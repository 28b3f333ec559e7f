"""featuresynth/experiment/melgan.py:11-46: MelGanGenerator + shared-weight three-scale
MelGanDiscriminator, hinge sub-losses (the Experiment defaults) -- BASELINE config 4."""
from ..audio.representation import RawAudio
from ..discriminator import MelGanDiscriminator
from ..generator.full import MelGanGenerator
from ..loss import mel_gan_disc_loss, mel_gan_gen_loss
from .experiment import Experiment
from .init import weights_init


class MultiScaleMelGanExperiment(Experiment):
    def __init__(self, **kw):
        total_samples = 8192
        samplerate = 22050
        n_fft, hop, n_mels = 1024, 256, 128
        feature_size = total_samples // hop
        super().__init__(
            generator=MelGanGenerator(feature_size, n_mels),
            discriminator=MelGanDiscriminator(),
            learning_rate=1e-4,
            feature_size=feature_size,
            audio_repr_class=RawAudio,
            generator_loss=mel_gan_gen_loss,
            discriminator_loss=mel_gan_disc_loss,
            g_init=weights_init,
            d_init=weights_init,
            feature_funcs={'audio': ('DeviceAudioStore.audio', (samplerate,)),
                           'spectrogram': ('Audio2Mel', (samplerate, n_fft, hop, n_mels))},
            total_samples=total_samples,
            feature_channels=n_mels,
            samplerate=samplerate,
            **kw)
