function [mu, sigma, alpha, AEPE, Energy, logP] = qgmap_chunked(variant, options, I1, I2)
%QGMAP_CHUNKED  The reference's while-loop (gqmap_gpu_mixture.m:26-76) driven from MATLAB over the stateful MEX
%   interface: one gqmap_mex('step',h,n) per monitoring period instead of ~40 gpuArray kernels per iteration.
its = options.its; every = 300; if isfield(options,'log_every'), every = options.log_every; end
AEPE = NaN(its,1); Energy = zeros(its,1); logP = NaN(its,1); best_aepe = Inf; mark = 1;
h = gqmap_mex('create', variant, options, double(I1), double(I2));
cleanup = onCleanup(@() gqmap_mex('destroy', h));
if isfield(options,'init'), S = options.init; S.T = options.temperature; gqmap_mex('set_state', h, S);
else, seed = 0; if isfield(options,'seed'), seed = options.seed; end; gqmap_mex('init_state', h, seed); end
it = 1; stopped = 0;
while ~stopped && it <= its
    if it == 1, nextmon = 1; else, nextmon = ceil(it/every)*every; end
    n = min(nextmon, its) - it + 1;
    [E, dmu, dsig, nit, stopped] = gqmap_mex('step', h, n, its);
    Energy(it:it+nit-1) = E(1:nit); it = it + nit; last = it - 1;
    if nit > 0 && (last == 1 || mod(last, every) == 0)
        map = gqmap_mex('map', h);
        if isfield(options,'trueFlow') && ~isempty(options.trueFlow)
            AEPE(last) = gqmap_mex('aepe', h, map, options.trueFlow, options.unknownIdx);
            best_aepe = min(best_aepe, AEPE(last));
        end
        logP(last) = gqmap_mex('logp', h, map); mark = last;
        if isfield(options,'dir') && ~isempty(options.dir)
            if variant == 1, flow = repelem(map,4,4); flc = flowToColor_mex(flow(5:end-4,5:end-4,:));
            else, flc = flowToColor_mex(map); end
            imwrite(flc, [options.dir, '/', num2str(last), '.png']);
        end
    end
    if nit > 0
        fprintf('[%3d], \x0394(mu) = %e, \x0394(sigma) = %e, Energy = %e, AEPE=%e,logP=%e \n', ...
            last, dmu(nit), dsig(nit), E(nit), best_aepe, logP(mark));
    end
    if nit < n, break; end
end
S = gqmap_mex('get_state', h);
mu = cat(4, S.muu, S.muv); sigma = cat(4, S.sigmau, S.sigmav); alpha = S.alpha;
end
