import sys; sys.path.insert(0,'em-spec_b200'); sys.path.insert(0,'.')
import torch, emspec, bench
for gate in (-65.0, -200.0):
    S=600*48000
    pcm=bench.synth_device(S,0,torch.device('cuda'))
    eng=emspec.Engine(n_fft=4096,hop=128,noise_gate_db=gate, flags=3|4)
    F=eng.frame_count(S)
    idx=torch.empty((1,F,2049),dtype=torch.uint8,device='cuda')
    for i in range(3):
        eng.process_grid(pcm,out=(None,idx))
        print(gate, 'points ms', eng.stage_ms(0), 'post ms', eng.stage_ms(2), 'nonzero frac', (idx!=0).float().mean().item() if i==2 else '')
    eng.close()
