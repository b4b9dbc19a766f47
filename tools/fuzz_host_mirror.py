"""Sanitizer run of the C++ host mirror's parsers (include/q3tts.hpp: JSON, config.json, safetensors header, WAV): builds
tests/cpp/host_mirror_check.cpp with -fsanitize=address,undefined and feeds it valid files, every truncation and ~900 randomly
damaged variants.  CPU only; scratch under gpurun_out/asan.  Last run (round 1): 0 findings.
    python tools/fuzz_host_mirror.py"""
import json, os, random, shutil, struct, subprocess, sys
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT)
import numpy as np, torch
from qwen3_tts_rs_b200 import formats as F, spec as S
d=os.path.join(ROOT,'gpurun_out','asan'); exe=os.path.join(d,'hmc')
os.makedirs(d,exist_ok=True)
libdir=os.path.join(ROOT,'qwen3_tts_rs_b200')
subprocess.run(['g++','-std=c++17','-O1','-g','-fsanitize=address,undefined','-fno-omit-frame-pointer','-I',os.path.join(ROOT,'include'),
                os.path.join(ROOT,'tests','cpp','host_mirror_check.cpp'),'-o',exe,'-L',libdir,'-lq3tts_b200','-Wl,-rpath,'+libdir],check=True)
env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
bad=0
def run(*a):
    global bad
    r=subprocess.run([exe,*map(str,a)],capture_output=True,text=True,errors="replace",env=env)
    if 'Sanitizer' in r.stderr or 'runtime error' in r.stderr or r.returncode<0 or r.returncode>=128:
        bad+=1; print('SANITIZER/CRASH', a, r.returncode, r.stderr[-800:])
    return r
rnd=random.Random(1)
# json: valid + every truncation + random byte damage
text=json.dumps({"a":[1,2.5,{"b":"x\"yé😀","c":None}],"d":{"e":[True,False]},"talker_config":{"hidden_size":2048}})
p=d+'/x.json'
for cut in range(len(text)+1):
    open(p,'w').write(text[:cut]); run('json',p); run('config',p)
for _ in range(300):
    g=bytearray(text.encode())
    for _ in range(rnd.randint(1,4)): g[rnd.randrange(len(g))]=rnd.randrange(256)
    open(p,'wb').write(bytes(g)); run('json',p); run('config',p)
# safetensors: valid, truncations of header region, random damage
st=d+'/x.safetensors'
F.save_safetensors({"talker.model.norm.weight":torch.randn(64).to(torch.bfloat16),"decoder.x":torch.randn(3,4),"h":torch.randn(5).half()}, st)
raw=open(st,'rb').read()
for cut in list(range(0,min(len(raw),400),3))+[len(raw)-1]:
    open(st,'wb').write(raw[:cut]); run('safetensors',st)
for _ in range(300):
    g=bytearray(raw)
    for _ in range(rnd.randint(1,4)): g[rnd.randrange(min(len(g),300))]=rnd.randrange(256)
    open(st,'wb').write(bytes(g)); run('safetensors',st)
# wav
w=d+'/x.wav'
F.save_wav(w, np.linspace(-1,1,100,dtype=np.float32), 24000); raw=open(w,'rb').read()
for cut in range(0,len(raw),5):
    open(w,'wb').write(raw[:cut]); run('wav',w)
for _ in range(300):
    g=bytearray(raw)
    for _ in range(rnd.randint(1,4)): g[rnd.randrange(min(len(g),60))]=rnd.randrange(256)
    open(w,'wb').write(bytes(g)); run('wav',w)
os.makedirs(d+'/f',exist_ok=True); run('formats',d+'/f'); run('prompts',151936,'1,2,3','4,5'); run('prompts',2048,'','')
print('sanitizer findings:',bad)
shutil.rmtree(d,ignore_errors=True)
sys.exit(1 if bad else 0)
