import sys, os, torch
sys.path.insert(0,'/root/repo/vae-diffusion-toy-crystals_b200'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import latent_prior_oracle as po, toycrystals_oracle as orc
from toycrystals_b200.models import diffusion_prior as pshim
G=torch.load('/root/repo/tests/golden/prior.pt',weights_only=False)
m=pshim.DiffusionPriorFiLM(**po.PRIOR_CFG,precision='bf16'); m.load_state_dict(po.prior_default_init(0)); m=m.cuda().eval()
for fw in G['forwards']:
    t=torch.full((G['n'],),fw['t'],dtype=torch.int64)
    e=m((G['z_init']*1.7).cuda(),t.cuda(),G['y_cat'].cuda(),G['y_cont'].cuda())
    print('t',fw['t'],'rel-L2',orc.rel_l2(e,fw['eps']))
s=pshim.DiffusionSchedule.linear(1000,1e-4,0.05,torch.device('cuda'))
z0,tr=s.ddim_sample(m,G['y_cat'].cuda(),G['y_cont'].cuda(),n_steps=50,z_init=G['z_init'].cuda(),return_trace=True)
sd=po.prior_default_init(0); worst=0
for k in range(tr.eps.shape[0]):
    t=G['timesteps'][k].repeat(G['n'])
    want=po.film_prior(sd,po.PRIOR_CFG,tr.z_in[k].cpu(),t,G['y_cat'],G['y_cont'])
    worst=max(worst,orc.rel_l2(tr.eps[k],want))
print('ddim teacher-forced worst',worst,'final z0 rel',orc.rel_l2(z0,G['z0']))
