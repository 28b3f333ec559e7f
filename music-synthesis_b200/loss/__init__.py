from .loss import (least_squares_generator_loss, hinge_generator_loss,  # noqa: F401
                   least_squares_disc_loss, hinge_discriminator_loss, mel_gan_disc_loss,
                   mel_gan_feature_loss, mel_gan_gen_loss)
