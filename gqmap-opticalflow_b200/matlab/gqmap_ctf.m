function [mu, sigma, rou, AEPE, Energy] = gqmap_ctf(options, I1, I2, GRDT)
%GQMAP_CTF  Drop-in for legacy/gqmap_ctf.m (the per-level solver of legacy/optical_flow_ctf.m) on libqgmap.so.
%   Same inputs and outputs (mu, sigma: M x N x 2; rou: M x N x 2 x 2; AEPE, Energy: its x 1).  The loop of
%   legacy/gqmap_ctf.m:22-65 runs in the CUDA iteration kernel with that file's constants: single Gaussian (L=1), constant
%   step 0.07 (:36), sigma stepped with step*0.3 and clamped to [0.01,25] (:48-49), correlations clamped to +-0.999 (:7,:50-51),
%   no entropy term, clamp range of the means = extrema of GRDT (:4), initial state as :14-20.
%   Differences (see DESIGN.md): the data term uses the live solver's exact bicubic instead of a nearest lookup into a 64x
%   upsampled second frame (:10,:76); AEPE is evaluated for the final beliefs only (the reference evaluates it every iteration).
[M, N] = size(I1);
o = options;
o.L = 1; o.temperature = 0; o.drate = 1;
o.minu = min(min(GRDT(:,:,1))); o.maxu = max(max(GRDT(:,:,1)));
o.minv = min(min(GRDT(:,:,2))); o.maxv = max(max(GRDT(:,:,2)));
o.step0 = 0.07; o.step_tau = Inf; o.sigma_step_scale = 0.3; o.sigma_min = 0.01; o.sigma_max = 25; o.corr_tor = 0.999;
o.init = struct('muu', o.minu + rand(M,N)*(o.maxu-o.minu), 'muv', o.minv + rand(M,N)*(o.maxv-o.minv), ...
                'sigmau', rand(M,N) + 3, 'sigmav', rand(M,N) + 3, 'pn', zeros(M,N), 'rou', zeros(M,N,1,2,2), 'w', 0);
h = gqmap_mex('create', 0, o, double(I1), double(I2));
cleaner = onCleanup(@() gqmap_mex('destroy', h));
gqmap_mex('set_state', h, o.init);
[E, ~, ~, nit] = gqmap_mex('step', h, options.its, options.its);
S = gqmap_mex('get_state', h);
mu = cat(3, S.muu, S.muv); sigma = cat(3, S.sigmau, S.sigmav); rou = reshape(S.rou, M, N, 2, 2);
Energy = zeros(options.its, 1); Energy(1:nit) = E(1:nit);
AEPE = nan(options.its, 1);
if size(GRDT,1) == M && size(GRDT,2) == N
    M_ = 2:M-1; N_ = 2:N-1;
    AEPE(max(nit,1)) = mean(mean(sqrt((GRDT(M_,N_,1)-mu(M_,N_,1)).^2 + (GRDT(M_,N_,2)-mu(M_,N_,2)).^2)));
end
end
