function results = qgmap_middlebury_demo(which, root)
% QGMAP_MIDDLEBURY_DEMO  run one of the reference's two experiment set-ups on top of the B200 path.
%   results = qgmap_middlebury_demo('full')   % the set-up of optical_flow.m:3-28      (Teddy, Cones; K=9, L=3)
%   results = qgmap_middlebury_demo('super')  % the set-up of optical_flowSuper.m:3-35 (five sequences; K=11, L=3, T=0.2)
% The reference's own driver scripts run unchanged against matlab/gqmap_gpu_mixture.m / gqmap_gpuSuper_mix_entropy.m and the Linux MEX
% files (that is the point of the drop-in); this function only shows the calls in one place.  `root` = folder holding middlebury/.
if nargin < 2, root = '.'; end
setups.full  = struct('solver', @gqmap_gpu_mixture,          'seqs', {{'Teddy','Cones'}}, ...
                      'K', 9,  'lambdas', 5,  'temperature', 0,   'drate', 0.5);
setups.super = struct('solver', @gqmap_gpuSuper_mix_entropy, 'seqs', {{'Venus','Hydrangea','Urban2','Urban3','Grove3'}}, ...
                      'K', 11, 'lambdas', 16, 'temperature', 0.2, 'drate', 0.75);
s = setups.(which);
results = struct('name', {}, 'AEPE', {}, 'Energy', {}, 'logP', {}, 'mu', {}, 'sigma', {}, 'alpha', {});
for k = 1:numel(s.seqs)
    seq = fullfile(root, 'middlebury', s.seqs{k});
    frames = cellfun(@(f) double(rgb2gray(imread(fullfile(seq, f)))), {'frame10.png','frame11.png'}, 'UniformOutput', false);
    o = struct('K', s.K, 'L', 3, 'its', 30000, 'epsn', 1e-6, 'lambdad', 1, 'lambdas', s.lambdas, ...
               'temperature', s.temperature, 'drate', s.drate);
    [~, o.trueFlow, o.minu, o.maxu, o.minv, o.maxv, o.unknownIdx] = flowToColor_mex(readFlowFile(fullfile(seq, 'flow10.flo')));
    r.name = s.seqs{k};
    [r.mu, r.sigma, r.alpha, r.AEPE, r.Energy, r.logP] = s.solver(o, frames{1}, frames{2});
    results(end+1) = orderfields(r, results); %#ok<AGROW>
end
end
