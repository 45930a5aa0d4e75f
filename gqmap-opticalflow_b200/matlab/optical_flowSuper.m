% optical_flowSuper.m -- thin example driver mirroring the reference's optical_flowSuper.m:3-35 on top of the B200 path.
clear;
testdata = {'Venus','Hydrangea','Urban2','Urban3','Grove3'};
for ti = 1:numel(testdata)
    name = testdata{ti};
    img1 = double(rgb2gray(imread(['middlebury/',name,'/frame10.png'])));
    img2 = double(rgb2gray(imread(['middlebury/',name,'/frame11.png'])));
    [gdt_img, options.trueFlow, options.minu, options.maxu, options.minv, options.maxv, options.unknownIdx] = ...
        flowToColor_mex(readFlowFile(['middlebury/',name,'/flow10.flo']));
    options.K = 11; options.its = 30000; options.epsn = 0.001^2; options.lambdas = 16; options.lambdad = 1;
    options.L = 3; options.temperature = 0.2; options.drate = 0.75;
    [mu, sigma, alpha, AEPE, Energy, logP] = gqmap_gpuSuper_mix_entropy(options, img1, img2);
    save([name,'_super.mat'], 'options', 'mu', 'sigma', 'alpha', 'AEPE', 'Energy', 'logP');
end
