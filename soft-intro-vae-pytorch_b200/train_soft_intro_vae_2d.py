"""
2-D toy Soft-IntroVAE (BASELINE config 1: "8Gaussians 2D SoftIntroVAE on CPU -- plumbing / correctness, no GPU").

Mirror of ``soft_intro_vae_2d/train_soft_intro_vae_2d.py`` of the reference: same public names and signatures
(ToyDataset :29-115, sample_2d_data :118-177, helpers :185-394, EncoderSimple / DecoderSimple / SoftIntroVAESimple
:402-483, train_soft_intro_vae_toy :486-725) and the same consumption of the numpy / python / torch random streams, so
a seeded run reproduces the reference's log (tests/golden/toy2d.pt).  This configuration is the reference's own
CPU-runnable case: a 0.27 M-parameter MLP on 512x2 batches.  It is deliberately NOT on the CUDA engine (SURVEY 8a
row a16: "plumbing only, no kernels"); the loss algebra is the one the fused CUDA loss pass implements, written once
in `_e_losses` / `_d_losses`.
"""
import os
import random
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim

_SQ2 = 1.0 / np.sqrt(2)
_RING8 = [(1, 0), (-1, 0), (0, 1), (0, -1), (_SQ2, _SQ2), (_SQ2, -_SQ2), (-_SQ2, _SQ2), (-_SQ2, -_SQ2)]


def _plt():
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
        return plt
    except ImportError:
        return None


class ToyDataset:
    """Infinite 2-D samplers: '8Gaussians', '25Gaussians', 'Sequential8Gaussians', '2spirals', 'checkerboard', 'rings'."""

    def __init__(self, distr='8Gaussians', dim=2, scale=2, iter_per_mode=100):
        self.distr, self.dim, self.scale = distr, dim, scale
        # 25-Gaussian table: 4000 sweeps over the 5x5 grid, one N(0, 0.05^2) pair per point (same draw order as the
        # reference's triple loop), shuffled, scaled by 1/2.828
        pts = np.random.randn(100000, 2) * 0.05
        grid = np.array([(2 * x, 2 * y) for x in range(-2, 3) for y in range(-2, 3)], dtype=np.float64)
        pts += np.tile(grid, (100000 // 25, 1))
        self.dataset = pts.astype('float32')
        np.random.shuffle(self.dataset)
        self.dataset /= 2.828
        self.range = 2 if distr == '25Gaussians' else 1
        self.curr_iter, self.curr_mode, self.iter_per_mode = 0, 0, iter_per_mode

    def _around(self, centers, batch_size, sig):
        noise = np.random.randn(batch_size, 2) * sig
        pts = noise + np.array(centers, dtype=np.float64)
        return torch.FloatTensor((pts.astype('float32')) / np.float32(1.414))

    def next_batch(self, batch_size=64, device=None, sig=0.02):
        if self.distr in ('2spirals', 'checkerboard', 'rings'):
            return sample_2d_data(self.distr, batch_size).to(device)
        centers = [(self.scale * x, self.scale * y) for x, y in _RING8]
        if self.distr == '8Gaussians':
            picks = []
            noise = np.empty((batch_size, 2))
            for i in range(batch_size):             # interleaved numpy / python draws, like the reference loop
                noise[i] = np.random.randn(2) * sig
                picks.append(random.choice(centers))
            pts = (noise + np.array(picks)).astype('float32') / np.float32(1.414)
            return torch.FloatTensor(pts).to(device)
        if self.distr == '25Gaussians':
            i = np.random.randint(100000 // batch_size)
            return torch.FloatTensor(self.dataset[i * batch_size:(i + 1) * batch_size]).to(device) * self.scale
        if self.distr == 'Sequential8Gaussians':
            out = self._around([centers[self.curr_mode]] * batch_size, batch_size, .02)
            if self.curr_iter % self.iter_per_mode == self.iter_per_mode - 1:
                self.curr_mode = (self.curr_mode + 1) % 8
            self.curr_iter += 1
            return out.to(device)
        return None


def sample_2d_data(dataset, n_samples):
    """'8gaussians' | '2spirals' | 'checkerboard' | 'rings' (the BNAF toy densities)"""
    z = torch.randn(n_samples, 2)
    if dataset == '8gaussians':
        centers = torch.tensor([(4 * x, 4 * y) for x, y in
                                [(1, 0), (-1, 0), (0, 1), (0, -1), (_SQ2, _SQ2), (-_SQ2, _SQ2), (_SQ2, -_SQ2), (-_SQ2, -_SQ2)]])
        return _SQ2 * (0.5 * z + centers[torch.randint(len(centers), size=(n_samples,))])
    if dataset == '2spirals':
        n = torch.sqrt(torch.rand(n_samples // 2)) * 540 * (2 * np.pi) / 360
        d1x = - torch.cos(n) * n + torch.rand(n_samples // 2) * 0.5
        d1y = torch.sin(n) * n + torch.rand(n_samples // 2) * 0.5
        x = torch.cat([torch.stack([d1x, d1y], dim=1), torch.stack([-d1x, -d1y], dim=1)], dim=0) / 3
        return x + 0.1 * z
    if dataset == 'checkerboard':
        x1 = torch.rand(n_samples) * 4 - 2
        x2_ = torch.rand(n_samples) - torch.randint(0, 2, (n_samples,), dtype=torch.float) * 2
        return torch.stack([x1, x2_ + x1.floor() % 2], dim=1) * 2
    if dataset == 'rings':
        q = n_samples // 4
        sizes = [q, q, q, n_samples - 3 * q]
        lins = [torch.linspace(0, 2 * np.pi, s + 1)[:-1] for s in sizes]
        # NB the reference pairs cos(linspace4) with sin(linspace3) for the 0.75 ring; kept
        xs = torch.cat([torch.cos(lins[0]), torch.cos(lins[0]) * 0.75, torch.cos(lins[2]) * 0.5, torch.cos(lins[3]) * 0.25])
        ys = torch.cat([torch.sin(lins[0]), torch.sin(lins[1]) * 0.75, torch.sin(lins[2]) * 0.5, torch.sin(lins[3]) * 0.25])
        x = torch.stack([xs, ys], dim=1) * 3.0
        x = x[torch.randint(0, n_samples, size=(n_samples,))]
        return x + torch.normal(mean=torch.zeros_like(x), std=0.08 * torch.ones_like(x))
    raise RuntimeError('Invalid `dataset` to sample from.')


# ---------------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------------
def reparameterize(mu, logvar):
    return torch.addcmul(mu, torch.randn_like(logvar), torch.exp(0.5 * logvar))


def load_model(model, pretrained):
    model.load_state_dict(torch.load(pretrained)['model'])


def save_checkpoint(model, epoch, iteration, prefix=""):
    os.makedirs("./saves/", exist_ok=True)
    path = "./saves/" + prefix + "model_epoch_{}_iter_{}.pth".format(epoch, iteration)
    torch.save({"epoch": epoch, "model": model.state_dict()}, path)
    print("model checkpoint saved @ {}".format(path))


def setup_grid(range_lim=4, n_pts=1000, device=torch.device("cpu")):
    x = torch.linspace(-range_lim, range_lim, n_pts)
    xx, yy = torch.meshgrid((x, x), indexing="ij")
    return xx, yy, torch.stack((xx.flatten(), yy.flatten()), dim=1).to(device)


def format_ax(ax, range_lim):
    ax.set_xlim(-range_lim, range_lim)
    ax.set_ylim(-range_lim, range_lim)
    ax.get_xaxis().set_visible(False)
    ax.get_yaxis().set_visible(False)
    ax.invert_yaxis()


def calc_reconstruction_loss(x, recon_x, loss_type='mse', reduction='sum'):
    recon_x, x = recon_x.view(x.size(0), -1), x.view(x.size(0), -1)
    if reduction not in ('sum', 'mean', 'none'):
        raise NotImplementedError
    if loss_type == 'mse':
        err = F.mse_loss(recon_x, x, reduction='none').sum(1)
        return err.sum() if reduction == 'sum' else err.mean() if reduction == 'mean' else err
    if loss_type == 'l1':
        return F.l1_loss(recon_x, x, reduction=reduction)
    if loss_type == 'bce':
        return F.binary_cross_entropy(recon_x, x, reduction=reduction)
    raise NotImplementedError


def calc_kl(logvar, mu, mu_o=10, is_outlier=False, reduce='sum'):
    inner = 1 + logvar - mu.pow(2) - logvar.exp()
    if is_outlier:
        inner = inner + 2 * mu * mu_o - mu_o.pow(2)
    kl = -0.5 * inner.sum(1)
    return kl.sum() if reduce == 'sum' else kl.mean() if reduce == 'mean' else kl


def _neg_elbo(model, x, beta_kl, beta_recon):
    mu, logvar, _, rec = model(x, deterministic=True)
    err = calc_reconstruction_loss(x, rec, loss_type='mse', reduction='none')
    while err.dim() > 1:
        err = err.sum(-1)
    return beta_kl * calc_kl(logvar=logvar, mu=mu, reduce="none") + beta_recon * err


def plot_vae_density(model, ax, test_grid, n_pts, batch_size, colorbar=False, beta_kl=1.0,
                     beta_recon=1.0, set_title=True, device=torch.device('cpu')):
    plt = _plt()
    model.eval()
    xx, yy, zz = test_grid
    with torch.no_grad():
        p_x = torch.cat([(-_neg_elbo(model, z.to(device), beta_kl, beta_recon)).exp() for z in zz.split(batch_size, dim=0)], 0)
    if plt is None:
        return p_x
    cmesh = ax.pcolormesh(xx.data.cpu().numpy(), yy.data.cpu().numpy(), p_x.view(n_pts, n_pts).data.cpu().numpy(), cmap=plt.cm.jet)
    ax.set_facecolor(plt.cm.jet(0.))
    if set_title:
        ax.set_title('VAE density')
    if colorbar:
        plt.colorbar(cmesh)
    return p_x


def plot_samples_density(dataset, model, scale, device):
    plt = _plt()
    model.eval()
    real = dataset.next_batch(batch_size=1024, device=device).data.cpu().numpy()
    fake = model.sample(torch.randn(size=(1024, model.zdim)).to(device)).data.cpu().numpy()
    if plt is None:
        return None
    fig = plt.figure(figsize=(18, 6))
    for i, (pts, title, col) in enumerate(((real, 'Real Data', None), (fake, 'Fake Samples', 'g'))):
        ax = fig.add_subplot(1, 3, i + 1)
        ax.scatter(pts[:, 0], pts[:, 1], s=8, c=col)
        ax.set_xlim((-scale * 2, scale * 2))
        ax.set_ylim((-scale * 2, scale * 2))
        ax.set_axis_off()
        ax.set_title(title)
    ax3 = fig.add_subplot(1, 3, 3)
    grid = setup_grid(range_lim=scale * 2, n_pts=1024, device=torch.device('cpu'))
    plot_vae_density(model, ax3, grid, n_pts=1024, batch_size=256, colorbar=False, beta_kl=1.0, beta_recon=1.0,
                     set_title=False, device=device)
    ax3.set_axis_off()
    ax3.set_title("Density Estimation")
    return fig


def calculate_elbo_with_grid(model, evalset, test_grid, beta_kl=1.0, beta_recon=1.0, batch_size=512, num_iter=100,
                             device=torch.device("cpu")):
    model.eval()
    _, _, zz = test_grid
    with torch.no_grad():
        grid = torch.cat([_neg_elbo(model, z.to(device), beta_kl, beta_recon) for z in zz.split(batch_size, dim=0)], 0)
        elbos = torch.cat([_neg_elbo(model, evalset.next_batch(batch_size=batch_size, device=device), beta_kl, beta_recon)
                           for _ in range(num_iter)], dim=0)
    return (elbos / torch.cat([grid, elbos], dim=0).sum()).mean().data.cpu().item()


def calculate_sample_kl(model, evalset, num_samples=5000, device=torch.device("cpu"), hist_bins=100, use_jsd=False,
                        xy_range=(-2, 2)):
    rng = [[xy_range[0], xy_range[1]], [xy_range[0], xy_range[1]]]
    real = evalset.next_batch(batch_size=num_samples, device=device).data.cpu().numpy()
    fake = model.sample_with_noise(num_samples=num_samples, device=device).data.cpu().numpy()
    rh = torch.tensor(np.histogram2d(real[:, 0], real[:, 1], bins=hist_bins, density=True, range=rng)[0]).to(device)
    fh = torch.tensor(np.histogram2d(fake[:, 0], fake[:, 1], bins=hist_bins, density=True, range=rng)[0]).to(device)
    if use_jsd:
        mid = 0.5 * (fh + rh)
        return (0.5 * (F.kl_div(torch.log(rh + 1e-14), mid, reduction='batchmean')
                       + F.kl_div(torch.log(fh + 1e-14), mid, reduction='batchmean'))).data.cpu().item()
    return F.kl_div(torch.log(fh + 1e-14), rh, reduction='batchmean').data.cpu().item()


# ---------------------------------------------------------------------------------------------------------------
# models
# ---------------------------------------------------------------------------------------------------------------
def _mlp(d_in, d_out, n_layers, num_hidden):
    main = nn.Sequential()
    main.add_module('input', nn.Linear(d_in, num_hidden))
    main.add_module('act0', nn.ReLU(True))
    for i in range(n_layers):
        main.add_module('hidden_%d' % (i + 1), nn.Linear(num_hidden, num_hidden))
        main.add_module('act_%d' % (i + 1), nn.ReLU(True))
    main.add_module('output', nn.Linear(num_hidden, d_out))
    return main


class EncoderSimple(nn.Module):
    def __init__(self, x_dim=2, zdim=2, n_layers=2, num_hidden=64):
        super().__init__()
        self.xdim, self.zdim, self.n_layer, self.num_hidden = x_dim, zdim, n_layers, num_hidden
        self.main = _mlp(x_dim, zdim * 2, n_layers, num_hidden)

    def forward(self, x):
        return self.main(x).view(x.size(0), -1).chunk(2, dim=1)


class DecoderSimple(nn.Module):
    def __init__(self, x_dim=2, zdim=2, n_layers=2, num_hidden=64):
        super().__init__()
        self.xdim, self.zdim, self.n_layer, self.num_hidden = x_dim, zdim, n_layers, num_hidden
        self.loggamma = nn.Parameter(torch.tensor(0.0))      # unused by the losses; part of the state_dict schema
        self.main = _mlp(zdim, x_dim, n_layers, num_hidden)

    def forward(self, z):
        return self.main(z.view(z.size(0), -1))


class SoftIntroVAESimple(nn.Module):
    def __init__(self, x_dim=2, zdim=2, n_layers=2, num_hidden=64):
        super().__init__()
        self.xdim, self.zdim, self.n_layer, self.num_hidden = x_dim, zdim, n_layers, num_hidden
        self.encoder = EncoderSimple(x_dim, zdim, n_layers, num_hidden)
        self.decoder = DecoderSimple(x_dim, zdim, n_layers, num_hidden)

    def forward(self, x, deterministic=False):
        mu, logvar = self.encode(x)
        z = mu if deterministic else reparameterize(mu, logvar)
        return mu, logvar, z, self.decode(z)

    def sample(self, z):
        return self.decode(z)

    def sample_with_noise(self, num_samples=1, device=torch.device("cpu")):
        return self.decode(torch.randn(num_samples, self.zdim).to(device))

    def encode(self, x):
        return self.encoder(x)

    def decode(self, z):
        return self.decoder(z)


# ---------------------------------------------------------------------------------------------------------------
# training
# ---------------------------------------------------------------------------------------------------------------
def _freeze(module, flag):
    for p in module.parameters():
        p.requires_grad = flag


def _e_losses(model, real, noise, s, beta_kl, beta_rec, beta_neg, loss_type):
    """E half (reference :560-613): returns lossE, z and the logged scalars.  Draw order: eps(real), eps(fake), eps(rec)."""
    fake = model.sample(noise)
    mu, logvar = model.encode(real)
    z = reparameterize(mu, logvar)
    rec = model.decoder(z)
    rec_det = model(real, deterministic=True)[3]
    l_rec = calc_reconstruction_loss(real, rec, loss_type=loss_type, reduction="mean")
    l_rec_det = calc_reconstruction_loss(real, rec_det.detach(), loss_type=loss_type, reduction="mean")
    kl_real = calc_kl(logvar, mu, reduce="mean")
    f_mu, f_lv, _, rec_fake = model(fake.detach())
    r_mu, r_lv, _, rec_rec = model(rec.detach())
    e_fake = (-2 * s * (beta_rec * calc_reconstruction_loss(fake, rec_fake, loss_type=loss_type, reduction="none")
                        + beta_neg * calc_kl(f_lv, f_mu, reduce="none"))).exp().mean()
    e_rec = (-2 * s * (beta_rec * calc_reconstruction_loss(rec, rec_rec, loss_type=loss_type, reduction="none")
                       + beta_neg * calc_kl(r_lv, r_mu, reduce="none"))).exp().mean()
    lossE = s * (beta_kl * kl_real + beta_rec * l_rec) + 0.25 * (e_fake + e_rec)
    return lossE, z, dict(rec_det=l_rec_det, kl_real=kl_real, e_rec=e_rec, e_fake=e_fake)


def _d_losses(model, real, noise, z, s, beta_kl, beta_rec, gamma_r, loss_type):
    """D half (reference :615-642).  Draw order: eps(rec), eps(fake)."""
    fake = model.sample(noise)
    rec = model.decoder(z.detach())
    l_rec = calc_reconstruction_loss(real, rec, loss_type=loss_type, reduction="mean")
    r_mu, r_lv = model.encode(rec)
    z_rec = reparameterize(r_mu, r_lv)
    f_mu, f_lv = model.encode(fake)
    z_fake = reparameterize(f_mu, f_lv)
    l_rr = calc_reconstruction_loss(rec.detach(), model.decode(z_rec.detach()), loss_type=loss_type, reduction="mean")
    l_rf = calc_reconstruction_loss(fake.detach(), model.decode(z_fake.detach()), loss_type=loss_type, reduction="mean")
    kl_fake, kl_rec = calc_kl(f_lv, f_mu, reduce="mean"), calc_kl(r_lv, r_mu, reduce="mean")
    lossD = s * (beta_rec * l_rec + 0.5 * beta_kl * (kl_fake + kl_rec) + gamma_r * 0.5 * beta_rec * (l_rr + l_rf))
    return lossD, dict(rec=l_rec, kl_fake=kl_fake, kl_rec=kl_rec)


def train_soft_intro_vae_toy(z_dim=2, lr_e=2e-4, lr_d=2e-4, batch_size=32, n_iter=30000, num_vae=0,
                             save_interval=1, recon_loss_type="mse", beta_kl=1.0, beta_rec=1.0,
                             beta_neg=1.0, test_iter=5000, seed=-1, pretrained=None, scale=1,
                             device=torch.device("cpu"), dataset="8Gaussians", gamma_r=1e-8):
    from tqdm import tqdm
    if seed != -1:
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        torch.backends.cudnn.deterministic = True
        print("random seed: ", seed)
    train_set = ToyDataset(distr=dataset)
    scale *= train_set.range
    model = SoftIntroVAESimple(x_dim=2, zdim=z_dim, n_layers=3, num_hidden=256).to(device)
    if pretrained is not None:
        load_model(model, pretrained)
    print(model)
    opt_e = optim.Adam(model.encoder.parameters(), lr=lr_e)
    opt_d = optim.Adam(model.decoder.parameters(), lr=lr_d)
    sch_e = optim.lr_scheduler.MultiStepLR(opt_e, milestones=(10000, 15000), gamma=0.1)
    sch_d = optim.lr_scheduler.MultiStepLR(opt_d, milestones=(10000, 15000), gamma=0.1)
    t0 = time.time()
    s = 0.5            # `dim_scale`, the normalising factor 's' of the paper
    plt = _plt()
    tag = "{}_bkl_{}_bneg_{}_brec_{}_seed_{}".format(dataset, beta_kl, beta_neg, beta_rec, seed)
    for it in tqdm(range(n_iter)):
        batch = train_set.next_batch(batch_size=batch_size, device=device)
        if it % save_interval == 0 and it > 0:
            save_checkpoint(model, (it // save_interval) * save_interval, it, '')
        model.train()
        if it < num_vae:
            real = batch.to(device)
            mu, logvar, _, rec = model(real)
            l_rec = calc_reconstruction_loss(real, rec, loss_type=recon_loss_type, reduction="mean")
            l_kl = calc_kl(logvar, mu, reduce="mean")
            opt_e.zero_grad()
            opt_d.zero_grad()
            (beta_rec * l_rec + beta_kl * l_kl).backward()
            opt_e.step()
            opt_d.step()
            if it % test_iter == 0:
                print("\nIter: {}/{} : time: {:4.4f}: ".format(it, n_iter, time.time() - t0)
                      + 'Rec: {:.4f}, KL: {:.4f} '.format(l_rec.data.cpu(), l_kl.data.cpu()))
        else:
            if batch.dim() == 3:
                batch = batch.unsqueeze(0)
            noise = torch.randn(size=(batch.size(0), z_dim)).to(device)
            real = batch.to(device)
            _freeze(model.encoder, True)
            _freeze(model.decoder, False)
            lossE, z, le = _e_losses(model, real, noise, s, beta_kl, beta_rec, beta_neg, recon_loss_type)
            opt_e.zero_grad()
            lossE.backward()
            opt_e.step()
            _freeze(model.encoder, False)
            _freeze(model.decoder, True)
            lossD, ld = _d_losses(model, real, noise, z, s, beta_kl, beta_rec, gamma_r, recon_loss_type)
            opt_d.zero_grad()
            lossD.backward()
            opt_d.step()
            if it % test_iter == 0:
                print("\nIter: {}/{} : time: {:4.4f}: ".format(it, n_iter, time.time() - t0)
                      + 'Rec: {:.4f} ({:.4f}), '.format(ld["rec"].data.cpu(), le["rec_det"].data.cpu())
                      + 'Kl_E: {:.4f}, expELBO_R: {:.4f}, expELBO_F: {:.4f}, '.format(le["kl_real"].data.cpu(), le["e_rec"].data.cpu(), le["e_fake"].cpu())
                      + 'Kl_F: {:.4f}, KL_R: {:.4f},'.format(ld["kl_fake"].data.cpu(), ld["kl_rec"].data.cpu())
                      + ' DIFF_Kl_F: {:.4f}'.format(-le["kl_real"].data.cpu() + ld["kl_fake"].data.cpu()))
            if torch.isnan(lossE) or torch.isnan(lossD):
                if plt is not None:
                    plt.close('all')
                raise SystemError("loss is NaN.")
        sch_e.step()
        sch_d.step()
        if (it % test_iter == 0 and it > 0) or it == n_iter - 1:
            print("\nplotting...")
            model.eval()
            fake = model.sample(torch.randn(size=(1024, z_dim)).to(device)).data.cpu().numpy()
            _scatter(plt, fake, scale, tag + "_iter_{}.png".format(it), 'g')
            if it == n_iter - 1:
                real_pts = train_set.next_batch(batch_size=1024, device=device).data.cpu().numpy()
                _scatter(plt, real_pts, scale, tag + "_iter_{}_real.png".format(it), None)
                print("plotting density...")
                grid = setup_grid(range_lim=scale * 2, n_pts=1024, device=torch.device('cpu'))
                if plt is not None:
                    fig, ax = plt.subplots(1, 1, figsize=(6, 6))
                    plot_vae_density(model, ax, grid, n_pts=1024, batch_size=256, colorbar=False, beta_kl=1.0, beta_recon=1.0,
                                     set_title=False, device=device)
                    ax.set_axis_off()
                    plt.savefig("density_" + tag + "_iter_{}.png".format(it), bbox_inches='tight')
                    plt.close()
            model.train()
    plot_samples_density(train_set, model, scale, device)
    res = {}
    print("estimating kl...")
    res['sample_kl'] = calculate_sample_kl(model, train_set, num_samples=5000, device=device, hist_bins=100, use_jsd=False,
                                           xy_range=(-2 * scale, 2 * scale))
    print("estimating jsd...")
    res['jsd'] = calculate_sample_kl(model, train_set, num_samples=5000, device=device, hist_bins=100, use_jsd=True,
                                     xy_range=(-2 * scale, 2 * scale))
    print("calculating elbo...")
    grid = setup_grid(range_lim=scale * 2, n_pts=1024, device=torch.device("cpu"))
    res['elbo'] = calculate_elbo_with_grid(model, train_set, test_grid=grid, beta_kl=1.0, beta_recon=1.0, device=device, batch_size=128)
    print("#" * 50)
    print("quantitative results:")
    print(f'dataset: {dataset}, beta_kl: {beta_kl}, beta_rec: {beta_rec}, beta_neg: {beta_neg}')
    print(f'grid-normalized elbo: {res["elbo"]:.4e}, kl: {res["sample_kl"]:.4f}, jsd: {res["jsd"]:.4f}')
    print("#" * 50)
    with open('./results_log_soft_intro_vae.txt', 'a') as fp:
        fp.writelines("{}_beta_kl_{}_beta_neg_{}_beta_rec_{}_gnelbo_{}_kl_{}_jsd_{}_seed_{}\n".format(
            dataset, beta_kl, beta_neg, beta_rec, res['elbo'], res['sample_kl'], res['jsd'], seed))
    return model


def _scatter(plt, pts, scale, fname, color):
    if plt is None:
        return
    fig, ax = plt.subplots(1, 1, figsize=(6, 6))
    ax.scatter(pts[:, 0], pts[:, 1], s=8, c=color)
    ax.set_xlim((-scale * 2, scale * 2))
    ax.set_ylim((-scale * 2, scale * 2))
    ax.set_axis_off()
    plt.savefig(fname, bbox_inches='tight')
    plt.close()


if __name__ == '__main__':
    hp = {'8Gaussians': (0.3, 0.9, 0.2), '2spirals': (0.5, 1.0, 0.2), 'checkerboard': (0.1, 0.2, 0.2), 'rings': (0.2, 1.0, 0.2)}
    ds = '8Gaussians'
    b_kl, b_neg, b_rec = hp[ds]
    train_soft_intro_vae_toy(z_dim=2, lr_e=2e-4, lr_d=2e-4, batch_size=512, n_iter=30_000, num_vae=2000, save_interval=5000,
                             recon_loss_type="mse", beta_kl=b_kl, beta_rec=b_rec, beta_neg=b_neg, test_iter=5000, seed=92,
                             scale=1 if ds == '8Gaussians' else 2, device=torch.device("cpu"), dataset=ds)
