import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import projected_langevin_sampling_b200 as b200
from test_gpu_parity import _problem, _build_pair, rel_err
for (n,d,m,j,ck,link) in [(777,3,33,131,"poisson","square"),(900,5,70,50,"multimodal","identity")]:
    x,y,z,ls,g = _problem(n,d,m,j,seed=n+j,cost_kind=ck,link=link)
    pls, orc = _build_pair(b200, x,y,z,ls,1.4,ck,link)
    m_k = orc.basis.approximation_dimension
    p = 0.5*torch.randn(m_k,j,generator=g,dtype=torch.float64)
    if ck=="poisson": p = p+0.3
    pc = p.cuda()
    F = pls.basis.calculate_untransformed_train_prediction_samples(pc)
    Fo = orc.basis.forward(p)
    print(ck, "F err", rel_err(F,Fo), "min|F|", Fo.abs().min().item())
    dc = pls.calculate_cost_derivative(pc); dco = orc.calculate_cost_derivative(p)
    print(" dc err", rel_err(dc,dco))
    # elementwise function on identical F
    dce = pls.cost.calculate_cost_derivative(Fo.cuda(), force_autograd=(ck=="multimodal"))
    print(" elementwise dc on same F", rel_err(dce, dco))
    el = (dce.cpu()-dco).abs()/dco.abs().clamp_min(1e-300)
    print(" max elementwise relative", el.max().item())
    print(" cost err", rel_err(pls.calculate_cost(pc), orc.calculate_cost(p)))
