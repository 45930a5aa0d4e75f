function [mu, sigma, alpha, AEPE, Energy, logP] = gqmap_gpu_mixture(options, I1, I2)
%GQMAP_GPU_MIXTURE  Full-resolution QGMAP inference on a B200 through libqgmap.so (no gpuArray / arrayfun).
%   Drop-in for the reference's gqmap_gpu_mixture(options,I1,I2): same options fields (trueFlow, unknownIdx, its, K, L,
%   temperature, drate, epsn, lambdad, lambdas, minu, maxu, minv, maxv), same six outputs with the same shapes
%   (mu,sigma: M x N x L x 2; alpha: 1 x 1 x L; AEPE,Energy,logP: its x 1, prefilled NaN/0/NaN).
%   New OPTIONAL fields: options.init (struct muu,muv,sigmau,sigmav,pn,rou,w), options.seed, options.alpha_mode
%   ('softmax' | 'projsplx'), options.device, options.devices (vector of CUDA ordinals: the frame pair is split into one row band
%   per GPU, boundary rows exchanged over NVLink by the library), options.log_every, options.verbose; options.dir, when present,
%   receives <it>.png at every monitored iteration exactly as in the reference (:59-62).
%   The whole loop (gradients, update, Energy, alpha update, MAP/AEPE/logP monitoring every 300 iterations) runs on the
%   GPU inside one MEX call; with options.verbose the loop is driven from MATLAB in chunks so progress can be printed.
if isfield(options, 'verbose') && options.verbose
    [mu, sigma, alpha, AEPE, Energy, logP] = qgmap_chunked(0, options, I1, I2);
else
    [mu, sigma, alpha, AEPE, Energy, logP] = gqmap_mex('solve', 0, options, double(I1), double(I2));
end
end
