% Build the MEX shims against libfftconv.so (replaces compile.m / cuda_compile.m of the reference).
% Run from this directory on a machine with MATLAB + the Parallel Computing Toolbox.
lib = fullfile('..', 'fftconv_b200');
inc = fullfile('..', '..', 'include');
names = {'cudaFFTData', 'cudaConvFFTData', 'cudaConvolutionFFT', 'cudaConvFFTDataStreams'};
for i = 1:numel(names)
  mex('-largeArrayDims', [names{i} '.cpp'], ['-I' inc], ['-L' lib], '-lfftconv', '-lmwgpu', '-lcudart', ...
      '-outdir', fullfile('..', '..', 'bin'));
end
