% replay.m — run the UNTOUCHED reference .m files on a golden fixture and compare.
%
% UNVERIFIED: neither MATLAB nor GNU Octave exists in the build container (SURVEY.md §8c),
% so this script has never been executed.  It is the intended way to pin the oracle against
% the real reference:
%
%   python tests/golden/export_mat.py        % writes tests/golden/<name>.mat (scipy.io.savemat)
%   cd <reference checkout>; addpath(<repo>/oracle)
%   replay('<repo>/tests/golden/ct16_perturbed.mat')
%
% For every hot-path function it prints the relative difference between the reference's
% outputs and the oracle outputs stored in the fixture (expected <= 1e-8 over the first
% iterations; see DESIGN.md "Parity").
function replay(matfile)
    g = load(matfile);
    if isfield(g, 'A_dense'), A = g.A_dense; B = g.B_dense;
    else
        A = sparse(double(g.A_ir) + 1, col_of(g.A_jc), g.A_pr, double(g.A_shape(1)), double(g.A_shape(2)));
        B = sparse(double(g.B_ir) + 1, col_of(g.B_jc), g.B_pr, double(g.B_shape(1)), double(g.B_shape(2)));
    end
    b = g.b(:); x_true = g.x_true(:); tol = g.tol; maxit = double(g.maxit); lam = g.lam;
    [x, e, r, it] = hybrid_ab_gmres_rtp(A, B, b, x_true, tol, maxit, lam);
    report('hybrid_ab_gmres_rtp', x, g.ab_rtp_x(:), r, g.ab_rtp_res(:), it, g.ab_rtp_it);
    [x, e, r, it] = hybrid_ba_gmres_rtp(A, B, b, x_true, tol, maxit, lam);
    report('hybrid_ba_gmres_rtp', x, g.ba_rtp_x(:), r, g.ba_rtp_res(:), it, g.ba_rtp_it);
    [x, e, r, it] = hybrid_lsqr_solver(A, b, x_true, tol, maxit, lam);
    report('hybrid_lsqr_solver', x, g.hybrid_lsqr_x(:), r, g.hybrid_lsqr_res(:), it, g.hybrid_lsqr_it);
    [x, e, r, it] = hybrid_lsmr_solver(A, b, x_true, tol, maxit, lam);
    report('hybrid_lsmr_solver', x, g.hybrid_lsmr_x(:), r, g.hybrid_lsmr_res(:), it, g.hybrid_lsmr_it);
    [x, e, r, it] = lsqr_solver(A, b, x_true, tol, maxit);
    report('lsqr_solver', x, g.lsqr_x(:), r, g.lsqr_res(:), it, g.lsqr_it);
    [x, e, r, ar, it] = lsmr_solver(A, b, x_true, tol, maxit);
    report('lsmr_solver', x, g.lsmr_x(:), r, g.lsmr_res(:), it, g.lsmr_it);
    lams = g.gcv_lams(:);
    types = {'ab', 'ba'};
    for t = 1:2
        v = arrayfun(@(l) gcv_function(l, A, B, b, size(A, 1), double(g.k_gcv), types{t}), lams);
        ref = g.(['gcv_' types{t} '_vals']);
        fprintf('gcv_function %s: max rel diff %.2e\n', types{t}, max(abs(v(:) - ref(:)) ./ abs(ref(:))));
    end
end

function c = col_of(jc)
    jc = double(jc(:)); n = numel(jc) - 1; c = zeros(jc(end), 1);
    for j = 1:n, c(jc(j) + 1:jc(j + 1)) = j; end
end

function report(name, x, xo, r, ro, it, ito)
    k = min(numel(r), numel(ro));
    fprintf('%s: iters %d vs %d, |x-x_o|/|x_o| = %.2e, max residual-history diff = %.2e\n', name, it, ito, ...
            norm(x - xo) / norm(xo), max(abs(r(1:k) - ro(1:k)) ./ abs(ro(1:k))));
end
