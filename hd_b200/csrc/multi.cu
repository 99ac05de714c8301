// multi.cu — single-process multi-GPU entry points (filled in by the next commit).
