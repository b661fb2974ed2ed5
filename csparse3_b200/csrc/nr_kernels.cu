// nr_kernels.cu -- placeholder, filled in below
