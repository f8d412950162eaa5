import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=4):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return reps * n / dt / 1e9
run(True, True, 1)
print("H2D only  GB/s", run(True, False)); print("D2H only  GB/s", run(False, True)); print("both, each GB/s", run(True, True))
import subprocess; print(subprocess.run("nvidia-smi -q | grep -A6 'GPU Link Info' | head -12; nproc; free -g | head -2", shell=True, capture_output=True, text=True).stdout)
