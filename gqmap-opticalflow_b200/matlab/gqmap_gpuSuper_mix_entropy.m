function [mu, sigma, alpha, AEPE, Energy, logP] = gqmap_gpuSuper_mix_entropy(options, I1, I2)
%GQMAP_GPUSUPER_MIX_ENTROPY  Super-pixel (4x4 block) QGMAP inference on a B200 through libqgmap.so.
%   Drop-in for the reference's gqmap_gpuSuper_mix_entropy(options,I1,I2): beliefs live on the Mo/4 x No/4 grid, the
%   temperature is annealed every 500 iterations (T = max(T*drate, 0.001)), sigma is clamped to [0.01, 25].
%   Same options and outputs as gqmap_gpu_mixture; see there for the optional fields.
if mod(size(I1,1),4) ~= 0 || mod(size(I1,2),4) ~= 0
    error('qgmap:arg', 'gqmap_gpuSuper_mix_entropy needs image sides divisible by 4.');
end
if isfield(options, 'verbose') && options.verbose
    [mu, sigma, alpha, AEPE, Energy, logP] = qgmap_chunked(1, options, I1, I2);
else
    [mu, sigma, alpha, AEPE, Energy, logP] = gqmap_mex('solve', 1, options, double(I1), double(I2));
end
end
