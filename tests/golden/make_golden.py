"""Generate golden vectors by executing the UNMODIFIED reference source
(/root/reference: models/svae.py, models/gmm.py, models/smm.py, distributions/*.py) over the
numpy TF-1.3-op shim in oracle/tf_shim, in float64.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs tests/golden/*.npz (committed).  Every random draw the reference makes through
tf.random_normal / tf.random_uniform / tf.multinomial / Dirichlet.sample is stored next to the
outputs so the oracle and the CUDA path can be fed the identical noise.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('VMP_REFERENCE', '/root/reference')


def import_reference():
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tf_shim'))
    sys.path.insert(0, REF)
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'tensorboard', 'tensorboard.backend',
                 'tensorboard.backend.event_processing',
                 'tensorboard.backend.event_processing.event_accumulator']:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['tensorboard.backend.event_processing.event_accumulator'].EventAccumulator = object
    sys.modules['matplotlib.colors'].ColorConverter = object
    sys.modules['matplotlib.colors'].ListedColormap = object
    import tensorflow as tf
    assert 'tf_shim' in tf.__file__
    from models import svae, gmm, smm, vae
    from distributions import gaussian, niw, dirichlet, student_t
    return tf, dict(svae=svae, gmm=gmm, smm=smm, vae=vae, gaussian=gaussian, niw=niw, dirichlet=dirichlet,
                    student_t=student_t)


def A(t):
    return np.asarray(t.a if hasattr(t, 'a') else t)


def take_log(tf):
    out = [(k, v.copy()) for k, v in tf.rng_log]
    del tf.rng_log[:]
    return out


def synth_encoder_outputs(rs, N, D, K, scale):
    """eta2_diag = -1/2 softplus(N(0,1)) (vae.py:40); eta1 = mu1 * (-2 eta2_diag), mu1 a K-cluster mixture."""
    eta2d = -0.5 * np.logaddexp(0.0, rs.randn(N, D))
    centres = scale * rs.randn(K, D)
    mu1 = centres[rs.randint(0, K, size=N)] + 0.5 * rs.randn(N, D)
    return mu1 * (-2.0 * eta2d), eta2d


def synth_decoder_outputs(seed, N, K, S, dobs, decoder):
    """Deterministic stand-in for (y, decoder outputs); tests regenerate it from the seed."""
    rs = np.random.RandomState(seed)
    y = rs.randn(N, dobs)
    if decoder == 'standard':
        return y, (rs.randn(N, K, S, dobs), np.exp(0.3 * rs.randn(N, K, S, dobs)))
    return np.sign(y), (np.zeros((N, K, S, dobs)), rs.randn(N, K, S, dobs))


def svae_case(tf, M, name, K, D, N, S, seed, rho, dobs=3, decoder='standard', keep_big=True, perturb=True,
              overlap=None):
    svae = M['svae']
    tf.reset_default_graph()
    del tf.rng_log[:]
    rs = np.random.RandomState(1000 + seed)
    prior, theta = svae.init_mm(K, D, seed=seed, theta_as_variable=True)
    log = take_log(tf)          # two random_uniform draws (prior: m_scale=0, theta: m_scale=5)
    uniform_theta = log[1][1]
    phi_gmm = svae.init_recognition_params(theta, K, seed=seed)
    normal_pi = take_log(tf)[0][1]
    if perturb:
        # move phi_gmm / theta away from their symmetric initial values so that every code path
        # (off-diagonal L, non-trivial A, b) is exercised
        mu_k, L_k, pi_k = phi_gmm
        mu_k.a[...] = mu_k.a + (0.3 if overlap is None else 0.1) * rs.randn(K, D)
        L_k.a[...] = L_k.a + (0.2 if overlap is None else 0.1 / D ** 0.5) * rs.randn(K, D, D)
        pi_k.a[...] = pi_k.a + 0.1 * rs.randn(K)
        xs0 = 2.0 * rs.randn(4 * K + 7, D)
        r0 = rs.dirichlet(np.ones(K), size=xs0.shape[0])
        star0 = svae.m_step(prior, tf.constant(xs0), tf.constant(r0))
        svae.update_gmm_params(theta, star0, 0.5)
    if overlap is not None:
        # overlapping recognition components (the D >= 32 cases above have one-hot responsibilities):
        # shrink the components' eta1 towards 0 so that their centres P2^-1 eta1 crowd together
        phi_gmm[0].a[...] = overlap * phi_gmm[0].a
    theta_before = [A(t).copy() for t in theta]
    phi_gmm_np = [A(t).copy() for t in phi_gmm]
    if overlap is None:
        eta1, eta2d = synth_encoder_outputs(rs, N, D, K, scale=2.0)
    else:
        _, eta2_phi2, _ = svae.unpack_recognition_gmm(phi_gmm)
        centres = np.linalg.solve(-2.0 * A(eta2_phi2), A(phi_gmm[0])[..., None])[..., 0]
        eta2d = -0.5 * np.logaddexp(0.0, rs.randn(N, D))
        mu1 = centres[rs.randint(0, K, size=N)] + 0.3 * rs.randn(N, D)
        eta1 = mu1 * (-2.0 * eta2d)
    x_k, log_r, phi_tilde, dbg = svae.e_step((tf.constant(eta1), tf.constant(eta2d)), phi_gmm, S, seed=seed)
    noise = take_log(tf)[0][1]                                  # [N,K,D,S]
    xs_all = svae.subsample_x(x_k, log_r, seed)
    u = take_log(tf)[0][1]                                      # [N,S]
    x_samples = xs_all[:, 0, :]
    yy, rec = synth_decoder_outputs(3000 + seed, N, K, S, dobs, decoder)
    elbo, details = svae.compute_elbo(tf.constant(yy), (tf.constant(rec[0]), tf.constant(rec[1])), theta,
                                      phi_tilde, x_k, log_r, decoder)
    star = svae.m_step(prior, x_samples, tf.exp(log_r))
    star_np = [A(t).copy() for t in star]
    svae.update_gmm_params(theta, star, rho)
    # noise / u / decoder outputs are NOT stored: they are regenerated in the tests from the seeds
    #   noise = RandomState(seed).standard_normal((N,K,D,S)); u = RandomState(seed).random_sample((N,S))
    assert np.array_equal(noise, np.random.RandomState(seed).standard_normal((N, K, D, S)))
    assert np.array_equal(u, np.random.RandomState(seed).random_sample((N, S)))
    out = dict(K=K, D=D, N=N, S=S, rho=rho, decoder=decoder, seed=seed, dobs=dobs, rec_seed=3000 + seed,
               uniform_theta=uniform_theta, normal_pi=normal_pi,
               eta1=eta1, eta2d=eta2d,
               log_r=A(log_r), x_samples=A(x_samples),
               elbo=A(elbo), details=np.array([A(d) for d in details]))
    for i, nm in enumerate(['alpha', 'A', 'b', 'beta', 'v_hat']):
        out['prior_' + nm] = A(prior[i])
        out['theta_' + nm] = theta_before[i]
        out['star_' + nm] = star_np[i]
        out['new_' + nm] = A(theta[i])
    for i, nm in enumerate(['mu_k', 'L_k', 'pi_k']):
        out['phi_' + nm] = phi_gmm_np[i]
    if keep_big:
        out.update(x_k=A(x_k), eta1_tilde=A(phi_tilde[0]), eta2_tilde=A(phi_tilde[1]),
                   w_eta1=A(dbg[0]), w_eta2=A(dbg[1]))
    else:
        out.update(x_k_every4=A(x_k)[::4])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    r = np.exp(A(log_r))
    print(name, 'elbo', float(A(elbo)), 'N_k', star_np[0][:4], 'mean entropy of r %.3f' % float(-(r * A(log_r)).sum(1).mean()),
          'rows with max r < 0.99: %.2f' % float((r.max(1) < 0.99).mean()))


def svae_smm_case(tf, M, name, K, D, N, S, seed, rho, dof):
    svae = M['svae']
    tf.reset_default_graph()
    del tf.rng_log[:]
    rs = np.random.RandomState(2000 + seed)
    # experiments.py:154-174 : theta = (alpha_k, mu_k, L_k, DoF)
    gmm_prior, theta0 = svae.init_mm(K, D, seed=seed, theta_as_variable=False)
    take_log(tf)
    with tf.variable_scope('theta'):
        mu_k, L_k = svae.make_loc_scale_variables(gmm_prior)
    phi_gmm = svae.init_recognition_params(theta0, K, seed=seed)
    take_log(tf)
    for t, s in ((mu_k, 1.0), (L_k, 0.2), (phi_gmm[0], 0.3), (phi_gmm[1], 0.2), (phi_gmm[2], 0.1)):
        t.a[...] = t.a + s * rs.randn(*t.a.shape)
    alpha = tf.Variable(A(theta0[0]) + rs.rand(K), name='alpha_k')
    dof_t = tf.constant(dof * np.ones(K))
    theta = (alpha, mu_k, L_k, dof_t)
    eta1, eta2d = synth_encoder_outputs(rs, N, D, K, scale=2.0)
    x_k, log_r, phi_tilde, dbg = svae.e_step((tf.constant(eta1), tf.constant(eta2d)), phi_gmm, S, seed=seed)
    noise = take_log(tf)[0][1]
    assert np.array_equal(noise, np.random.RandomState(seed).standard_normal((N, K, D, S)))
    y, rec = synth_decoder_outputs(3000 + seed, N, K, S, 2, 'standard')
    elbo, details = svae.compute_elbo_smm(tf.constant(y), (tf.constant(rec[0]), tf.constant(rec[1])), theta,
                                          phi_tilde, x_k, log_r, 'standard')
    alpha_before = A(alpha).copy()
    alpha_star = svae.m_step_smm(smm_prior=[gmm_prior[0]], r_nk=tf.exp(log_r))
    svae.update_gmm_params([theta[0]], [alpha_star], rho)
    mu_t, sigma_t = svae.unpack_smm((mu_k, L_k))
    np.savez_compressed(
        os.path.join(HERE, name + '.npz'), K=K, D=D, N=N, S=S, rho=rho, seed=seed, dobs=2, rec_seed=3000 + seed,
        eta1=eta1, eta2d=eta2d,
        phi_mu_k=A(phi_gmm[0]), phi_L_k=A(phi_gmm[1]), phi_pi_k=A(phi_gmm[2]),
        prior_alpha=A(gmm_prior[0]), theta_alpha=alpha_before, theta_mu=A(mu_k), theta_L=A(L_k), theta_dof=A(dof_t),
        unpacked_sigma=A(sigma_t), log_r=A(log_r), x_k=A(x_k), elbo=A(elbo),
        details=np.array([A(d) for d in details]), alpha_star=A(alpha_star), alpha_new=A(alpha))
    print(name, 'elbo', float(A(elbo)))


def mixture_cases(tf, M):
    gmm, smm = M['gmm'], M['smm']
    rs = np.random.RandomState(7)
    for name, N, K, D in (('gmm_sweep_a', 200, 5, 3), ('gmm_sweep_b', 97, 10, 6)):
        tf.reset_default_graph()
        del tf.rng_log[:]
        centres = 3.0 * rs.randn(K, D)
        x = centres[rs.randint(0, K, size=N)] + rs.randn(N, D)
        step, log_r, theta, (x_k, S_k, pi) = gmm.inference(tf.constant(x), K, seed=0)
        log = take_log(tf)
        r0 = [v for k, v in log if k == 'dirichlet'][0]
        # second sweep from the updated state (step is the assigned variable)
        mask = rs.rand(N, D) < 0.2
        alpha_k, beta_k, m_k, C_k, v_k = theta
        P_k = tf.matrix_inverse(C_k)
        r_miss, pi_miss = gmm.e_step_missing_data(tf.constant(x), alpha_k, beta_k, m_k, P_k, v_k, tf.constant(mask))
        np.savez_compressed(os.path.join(HERE, name + '.npz'), x=x, K=K, r0=r0, r_new=A(step), log_r=A(log_r),
                            alpha_k=A(alpha_k), beta_k=A(beta_k), m_k=A(m_k), C_k=A(C_k), v_k=A(v_k),
                            x_k=A(x_k), S_k=A(S_k), pi=A(pi), mask=mask, r_miss=A(r_miss), pi_miss=A(pi_miss))
        print(name, A(pi)[:3])
    for name, N, K, D, kappa in (('smm_sweep_a', 150, 4, 2, 5.0), ('smm_sweep_b', 120, 7, 8, 9999.0)):
        tf.reset_default_graph()
        del tf.rng_log[:]
        centres = 3.0 * rs.randn(K, D)
        x = centres[rs.randint(0, K, size=N)] + rs.randn(N, D)
        x[: N // 20] = 10 * rs.rand(N // 20, D) - 5     # outliers
        # sweep 1 from (r0 ~ Dirichlet, u0 = 1), as smm.inference builds it
        step, log_r, theta, (x_k, S_k, pi) = smm.inference(tf.constant(x), K, kappa, seed=0)
        log = take_log(tf)
        r0 = [v for k, v in log if k == 'dirichlet'][0]
        alpha_k, beta_k, m_k, C_k, v_k, kappa_k = theta
        P_k = tf.matrix_inverse(C_k)
        r1, u1, _ = smm.e_step(tf.constant(x), alpha_k, beta_k, m_k, P_k, v_k, kappa_k)
        # sweep 2: m_step with (r1,u1) then e_step, exercising u != 1
        prior = M['svae'].init_mm_params(K, D, alpha_scale=0.05 / K, beta_scale=0.5, m_scale=0, C_scale=D + 0.5,
                                         v_init=D + 0.5, seed=0, as_variables=False)
        take_log(tf)
        beta_0, m_0, C_0, v_0 = M['niw'].natural_to_standard(*prior[1:])
        alpha_0 = M['dirichlet'].natural_to_standard(prior[0])
        th2 = smm.m_step(tf.constant(x), r1, u1, alpha_0, beta_0, m_0, C_0, v_0)
        P2 = tf.matrix_inverse(th2[3])
        r2, u2, pi2 = smm.e_step(tf.constant(x), th2[0], th2[1], th2[2], P2, th2[4], kappa_k)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), x=x, K=K, kappa=kappa, r0=r0,
                            r1=A(r1), u1=A(u1), log_r=A(log_r), pi=A(pi),
                            alpha_k=A(alpha_k), beta_k=A(beta_k), m_k=A(m_k), C_k=A(C_k), v_k=A(v_k),
                            x_k=A(x_k), S_k=A(S_k),
                            alpha2=A(th2[0]), beta2=A(th2[1]), m2=A(th2[2]), C2=A(th2[3]), v2=A(th2[4]),
                            xk2=A(th2[5]), Sk2=A(th2[6]), r2=A(r2), u2=A(u2), pi2=A(pi2))
        print(name, A(pi2)[:3])


def distribution_cases(tf, M):
    gaussian, niw, dirichlet, student_t = M['gaussian'], M['niw'], M['dirichlet'], M['student_t']
    rs = np.random.RandomState(11)
    K, D, N, S = 4, 5, 6, 3

    def spd(*lead):
        a = rs.randn(*lead, D, D)
        return a @ np.swapaxes(a, -1, -2) + D * np.eye(D)
    mu, sigma = rs.randn(K, D), spd(K)
    e1, e2 = gaussian.standard_to_natural(tf.constant(mu), tf.constant(sigma))
    mu_b, sigma_b = gaussian.natural_to_standard(e1, e2)
    x = rs.randn(N, D)
    eta1_nk, eta2_nk = rs.randn(N, K, D), -0.5 * spd(N, K)
    w = rs.dirichlet(np.ones(K))
    lp = gaussian.log_probability_nat(tf.constant(x), tf.constant(eta1_nk), tf.constant(eta2_nk), tf.constant(w))
    lp_now = gaussian.log_probability_nat(tf.constant(x), tf.constant(eta1_nk), tf.constant(eta2_nk))
    xs = rs.randn(N, K, S, D)
    lps = gaussian.log_probability_nat_per_samp(tf.constant(xs), tf.constant(eta1_nk), tf.constant(eta2_nk))
    beta, m, C, v = rs.rand(K) + 0.5, rs.randn(K, D), spd(K), rs.rand(K) * 3 + D + 1
    A_, b_, beta_, vh_ = niw.standard_to_natural(tf.constant(beta), tf.constant(m), tf.constant(C), tf.constant(v))
    back = niw.natural_to_standard(A_, b_, beta_, vh_)
    em, eC = niw.expected_values(back)
    alpha = rs.rand(K) * 4 + 0.1
    elp = dirichlet.expected_log_pi(tf.constant(alpha))
    dof = rs.rand(K) * 8 + 1.5
    st = student_t.log_probability_per_samp(tf.constant(xs), tf.constant(mu), tf.constant(sigma), tf.constant(dof))
    stm = student_t.logprob_smm_mixture(tf.constant(x), tf.constant(mu), tf.constant(sigma), tf.constant(dof),
                                        tf.constant(np.log(w)))
    np.savez_compressed(os.path.join(HERE, 'distributions.npz'), mu=mu, sigma=sigma, eta1=A(e1), eta2=A(e2),
                        mu_back=A(mu_b), sigma_back=A(sigma_b), x=x, eta1_nk=eta1_nk, eta2_nk=eta2_nk, w=w,
                        logprob_nat=A(lp), logprob_nat_noweights=A(lp_now), xs=xs, logprob_per_samp=A(lps),
                        beta=beta, m=m, C=C, v=v, A=A(A_), b=A(b_), v_hat=A(vh_),
                        back_m=A(back[1]), back_C=A(back[2]), back_v=A(back[3]), exp_m=A(em), exp_C=A(eC),
                        alpha=alpha, expected_log_pi=A(elp), dof=dof, student_per_samp=A(st), student_mixture=A(stm))
    print('distributions ok')


def losses_inputs(seed, N, K, S, D, C=4):
    """Seeded inputs of the losses fixtures; tests regenerate them (nothing large is stored)."""
    rs = np.random.RandomState(seed)
    y = rs.randn(N, D)
    pred = y[:, None, None, :] + 0.7 * rs.randn(N, K, S, D)
    var = np.exp(0.4 * rs.randn(N, K, S, D))
    logits = 1.5 * rs.randn(N, K, S, D)
    r = rs.dirichlet(0.5 * np.ones(K), N)
    labels = rs.randint(0, C, N)
    return y, pred, var, logits, r, labels


def imputation_stub(N, K, S, D):
    """Deterministic stand-in for svae.inference inside imputation_losses: call c -> (means, vars/logits, log r)."""
    cnt = [0]

    def imp(y_pert):
        rs = np.random.RandomState(500 + cnt[0])
        cnt[0] += 1
        means = y_pert[:, None, None, :] + 0.3 * rs.randn(N, K, S, D)
        return means, np.exp(0.3 * rs.randn(N, K, S, D)), np.log(rs.dirichlet(np.ones(K), N))
    return imp


def losses_cases(tf, M):
    import importlib
    losses = importlib.import_module('losses')
    N, K, S, D, P = 23, 4, 5, 6, 3
    y, pred, var, logits, r, labels = losses_inputs(77, N, K, S, D)
    c = tf.constant
    yb = np.sign(y)
    out = dict(N=N, K=K, S=S, D=D, P=P, seed=77)
    out['weighted_mse'] = A(losses.weighted_mse(c(y), c(pred), c(r)))
    out['gauss_logprob'] = A(losses.diagonal_gaussian_logprob(c(y), c(pred), c(var), c(np.log(r))))
    lw3 = np.log(r)[:, :, None] + 0.1 * np.random.RandomState(3).randn(N, K, S)
    out['gauss_logprob_nks'] = A(losses.diagonal_gaussian_logprob(c(y), c(pred), c(var), c(lw3)))
    mask = A(losses.generate_missing_data_mask(c(y), 0.3, seed=0))
    out['mask'] = mask
    out['gauss_logprob_mask'] = A(losses.diagonal_gaussian_logprob(c(y), c(pred), c(var), c(np.log(r)), mask=c(mask)))
    out['bernoulli_logprob'] = A(losses.bernoulli_logprob(c(yb), c(logits), c(np.log(r))))
    out['bernoulli_logprob_mask'] = A(losses.bernoulli_logprob(c(yb), c(logits), c(np.log(r)), c(mask)))
    out['imputation_mse'] = A(losses.imputation_mse(c(y), c(pred), c(r), c(mask)))
    ent, pur = losses.purity(c(r), c(np.eye(4)[labels]))
    out['entropy'], out['purity'] = A(ent), A(pur)
    for dt, yy in (('standard', y), ('bernoulli', yb)):
        del tf.rng_log[:]
        stub = imputation_stub(N, K, S, D)
        mse, ll = losses.imputation_losses(c(yy), c(mask), lambda yp: tuple(c(t) for t in stub(A(yp))), P, S, seed=0,
                                           decoder_type=dt)
        log = take_log(tf)
        assert len(log) == P and all(k == 'random_normal' for k, _ in log)   # perturb_data is always Gaussian noise
        out['imp_noise'] = np.stack([v for _, v in log])
        out['imp_mse_' + dt], out['imp_ll_' + dt] = A(mse), A(ll)
    np.savez_compressed(os.path.join(HERE, 'losses.npz'), **out)
    print('losses', {k: float(v) for k, v in out.items() if np.ndim(v) == 0 and k not in 'NKSDP'})


if __name__ == '__main__':
    tf, M = import_reference()
    tf.set_float(np.float64)
    if '--losses-only' in sys.argv:
        losses_cases(tf, M)
        sys.exit(0)
    if '--overlap-only' not in sys.argv:
        main_cases = True
    else:
        main_cases = False
    # non-degenerate responsibilities at D = 32 / 64 (round 2)
    svae_case(tf, M, 'svae_d64_overlap', K=8, D=64, N=24, S=1, seed=6, rho=0.2, keep_big=False, overlap=0.2)
    svae_case(tf, M, 'svae_d32_overlap', K=16, D=32, N=32, S=2, seed=7, rho=0.2, keep_big=False, overlap=0.3)
    if not main_cases:
        sys.exit(0)
    svae_case(tf, M, 'svae_c1', K=10, D=2, N=100, S=10, seed=0, rho=0.1)
    svae_case(tf, M, 'svae_c2', K=10, D=6, N=274, S=10, seed=0, rho=0.2, dobs=6, keep_big=False)
    svae_case(tf, M, 'svae_init', K=10, D=2, N=37, S=2, seed=3, rho=0.1, perturb=False)
    svae_case(tf, M, 'svae_d8', K=5, D=8, N=33, S=3, seed=1, rho=0.2, decoder='bernoulli')
    svae_case(tf, M, 'svae_d16', K=7, D=16, N=21, S=2, seed=4, rho=0.2, keep_big=False)
    svae_case(tf, M, 'svae_d32', K=6, D=32, N=19, S=1, seed=2, rho=0.2, keep_big=False)
    svae_case(tf, M, 'svae_d64', K=4, D=64, N=9, S=1, seed=5, rho=0.2, keep_big=False)
    svae_smm_case(tf, M, 'svae_smm', K=6, D=3, N=41, S=4, seed=0, rho=0.1, dof=5.0)
    mixture_cases(tf, M)
    distribution_cases(tf, M)
    losses_cases(tf, M)
