"""
Drop-in replacement of ``soft_intro_vae_bootstrap/train_soft_intro_vae_bootstrap.py`` on the B200 engine.

Differences to the standard trainer (reference bootstrap file :192-217, :241-246, :576-581, :593-594, :618-623,
:635-641, :680-682): a third network ``target_decoder`` (same architecture, own random init, never optimised)
produces rec_rec / rec_fake in both half-steps; nothing is detached in the D half, so the decoder also receives the
gradient that flows loss -> target decoder (dgrad only) -> encoder (dgrad only) -> rec / fake; ``gamma_r`` defaults
to 1.0; every ``copy_to_target_freq`` epochs the target decoder is overwritten with the decoder's state_dict.
Inside the engine this is ``variant = 1`` of the same step graphs (csrc/engine.cu).
"""
import importlib
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))
_PKG = os.path.basename(_HERE)
_B = importlib.import_module(_PKG + ".train_soft_intro_vae")
_L = importlib.import_module(_PKG + ".lib")
if os.environ.get("SIVAE_ANNOUNCE") == "1":      # which file served `import train_soft_intro_vae_bootstrap` (tests/test_gpu_boundary.py)
    print("SIVAE_DROPIN %s %s" % (__name__, os.path.abspath(__file__)), file=sys.stderr)

ResidualBlock, Encoder, Decoder = _B.ResidualBlock, _B.Encoder, _B.Decoder
calc_kl, reparameterize, calc_reconstruction_loss = _B.calc_kl, _B.reparameterize, _B.calc_reconstruction_loss
str_to_list, is_image_file, record_scalar, record_image = _B.str_to_list, _B.is_image_file, _B.record_scalar, _B.record_image
load_model, save_checkpoint = _B.load_model, _B.save_checkpoint
introspective_iteration, vae_iteration = _B.introspective_iteration, _B.vae_iteration
DevicePrefetcher = _B.DevicePrefetcher


class SoftIntroVAE(_B.SoftIntroVAE):
    """SoftIntroVAE with a frozen target decoder (reference bootstrap file :172-246)."""
    _bootstrap = True

    def __init__(self, cdim=3, zdim=512, channels=(64, 128, 256, 512, 512, 512), image_size=256, conditional=False,
                 cond_dim=10):
        super().__init__(cdim, zdim, channels, image_size, conditional=conditional, cond_dim=cond_dim)
        # constructed after the decoder: same RNG consumption order as the reference
        self.target_decoder = Decoder(cdim, zdim, channels, image_size, conditional=conditional,
                                      conv_input_size=self.encoder.conv_output_size, cond_dim=cond_dim)
        self._wire()

    def forward(self, x, o_cond=None, deterministic=False, target=True):
        mu, logvar = self.encode(x, o_cond=o_cond)
        z = mu if deterministic else reparameterize(mu, logvar)
        y = self.decode_target(z, y_cond=o_cond) if target else self.decode(z, y_cond=o_cond)
        return mu, logvar, z, y

    def decode_target(self, z, y_cond=None):
        return self.target_decoder(z, y_cond=y_cond)


def train_soft_intro_vae(dataset='cifar10', z_dim=128, lr_e=2e-4, lr_d=2e-4, batch_size=128, num_workers=4,
                         start_epoch=0, exit_on_negative_diff=False, copy_to_target_freq=1,
                         num_epochs=250, num_vae=0, save_interval=50, recon_loss_type="mse",
                         beta_kl=1.0, beta_rec=1.0, beta_neg=1.0, test_iter=1000, seed=-1, pretrained=None,
                         device=torch.device("cpu"), num_row=8, gamma_r=1.0, with_fid=False):
    """Bootstrap trainer: same arguments as the reference's (adds copy_to_target_freq, gamma_r defaults to 1.0)."""
    return _B._run_training(SoftIntroVAE, int(copy_to_target_freq), dataset, z_dim, lr_e, lr_d, batch_size, num_workers,
                            start_epoch, exit_on_negative_diff, num_epochs, num_vae, save_interval, recon_loss_type,
                            beta_kl, beta_rec, beta_neg, test_iter, seed, pretrained, device, num_row, gamma_r, with_fid)


if __name__ == '__main__':
    try:
        train_soft_intro_vae(dataset="synthetic32", z_dim=128, batch_size=32, num_workers=0, num_epochs=2, num_vae=0,
                             beta_kl=1.0, beta_neg=256, beta_rec=1.0, device=torch.device("cuda:0"), save_interval=50,
                             start_epoch=0, lr_e=2e-4, lr_d=2e-4, pretrained=None, copy_to_target_freq=1, test_iter=1000,
                             with_fid=False)
    except SystemError:
        print("Error, probably loss is NaN, try again...")
