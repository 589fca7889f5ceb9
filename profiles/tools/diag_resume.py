import importlib, os, sys, tempfile, torch
sys.path.insert(0, os.getcwd())
PKG = "soft-intro-vae-pytorch_b200"
M = importlib.import_module(PKG + ".train_soft_intro_vae")
L = importlib.import_module(PKG + ".lib")
orig = M.introspective_iteration
log = []
def wrapped(model, real, noise, eps, hp, lr_e, lr_d, **kw):
    eng = model._ensure_engine(real.size(0))
    pre = (float(eng.mem[0].params.double().sum()), float(eng.mem[0].m.double().sum()), float(eng.mem[0].v.double().sum()),
           float(eng.mem[1].params.double().sum()), float(eng.mem[1].m.double().sum()),
           int(L.load().sivae_adam_get_step(eng.handle, 0)), int(L.load().sivae_adam_get_step(eng.handle, 1)))
    st = orig(model, real, noise, eps, hp, lr_e, lr_d, **kw)
    torch.cuda.synchronize()
    log.append(dict(real=float(real.double().sum()), noise=float(noise.double().sum()), eps=float(eps.double().sum()), pre=pre,
                    post=float(eng.mem[0].params.double().sum()), stats=[float(x) for x in st[:6]]))
    return st
M.introspective_iteration = wrapped
kw = dict(dataset="synthetic32:48", z_dim=32, batch_size=16, num_workers=0, num_vae=0, beta_kl=1.0, beta_neg=256,
          beta_rec=1.0, device=torch.device("cuda:0"), save_interval=1, lr_e=2e-4, lr_d=2e-4, seed=5, test_iter=1000, with_fid=False)
tmp = tempfile.mkdtemp()
a, b = os.path.join(tmp, "a"), os.path.join(tmp, "b")
os.makedirs(a); os.makedirs(b)
so = sys.stdout
sys.stdout = open(os.devnull, "w")
os.chdir(a)
M.train_soft_intro_vae(num_epochs=2, start_epoch=0, pretrained=None, **kw)
la = list(log); log.clear()
mid = [f for f in os.listdir("saves") if f.endswith("iter_3.pth")][0]
os.chdir(b)
M.train_soft_intro_vae(num_epochs=2, start_epoch=1, pretrained=os.path.join(a, "saves", mid), **kw)
lb = list(log)
sys.stdout = so
for i, (x, y) in enumerate(zip(la[3:], lb)):
    print("iter", i + 3)
    for k in x:
        print("   ", k, "SAME" if x[k] == y[k] else "DIFF", x[k] if x[k] != y[k] else "", y[k] if x[k] != y[k] else "")
