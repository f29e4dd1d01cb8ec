"""SURVEY 8d config 4's synthetic stand-in: 3-D ACM flow past a translating sphere with volume penalization, adaptive every step
(CDF44, threshold_mask, force_maxlevel_dealiasing), shared by the oracle run and the GPU run of tests/test_gpu_sphere3d.py.  No reference
fixture exists for a 3-D adaptive run with penalization (the reference's 3-D penalization fixture needs the insect module): parity is pinned
through the oracle, whose adaptive loop and penalization term are pinned by the 2-D fixtures."""
BS, G = 16, 6
INI = dict(dim=3, Bs=(BS, BS, BS), g=G, g_rhs=2, n_eqn=4, domain=(1.0, 1.0, 1.0), Jmax=4, discretization="FD_4th_central", penalization=True,
           use_sponge=False, c0=10.0, nu=1.0e-3, gamma_p=1.0, C_eta=1.0e-3, u_mean_set=(1.0, 0.0, 0.0), CFL=1.0, CFL_eta=0.99,
           time_max=1.0, write_method="fixed_time", write_time=1.0)
EPS, JMIN = 1.0e-3, 1
SPHERE = dict(center=(0.3, 0.28, 0.33), radius=0.08, velocity=(0.6, 0.2, -0.1))
