% optical_flow.m -- thin example driver mirroring the reference's optical_flow.m:3-28 on top of the B200 path.
clear;
testdata = {'Teddy','Cones'};
for ti = 1:numel(testdata)
    name = testdata{ti};
    img_1 = double(rgb2gray(imread(['middlebury/',name,'/frame10.png'])));
    img_2 = double(rgb2gray(imread(['middlebury/',name,'/frame11.png'])));
    [gdt_img, options.trueFlow, options.minu, options.maxu, options.minv, options.maxv, options.unknownIdx] = ...
        flowToColor_mex(readFlowFile(['middlebury/',name,'/flow10.flo']));
    options.K = 9; options.its = 30000; options.epsn = 0.001^2; options.lambdas = 5; options.lambdad = 1;
    options.L = 3; options.temperature = 0; options.drate = 0.5;
    [mu, sigma, alpha, AEPE, Energy, logP] = gqmap_gpu_mixture(options, img_1, img_2);
    save([name,'.mat'], 'options', 'AEPE', 'mu', 'sigma', 'alpha', 'Energy', 'logP');
end
