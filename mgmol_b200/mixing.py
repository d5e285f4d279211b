"""AndersonMix<T> (src/AndersonMix.h:21-52, src/AndersonMix.cc:27-319): Anderson
extrapolation of the trial solution from the last m residuals, the accelerator
ABPG::update_states applies to the preconditioned residual
(src/ABPG.cc:118-127).  Host control flow and an m x m solve; everything
grid-sized is delegated to the vector type (on the GPU: Orbitals, i.e. the
C-ABI BLAS-1 kernels), which must provide

    assign(y)   self -= y   dotProduct(y) -> float   axpy(a, y)   scal(a)

and `clone(x)` must return a new vector shaped like x."""
import numpy as np

MIN_DET_MAT = 0.01   # src/AndersonMix.cc:23-25
MAX_THETA = 0.5
MIN_THETA = -3.0


class AndersonMix:
    def __init__(self, m, beta, x, clone):
        self.m_, self.mm_, self.beta_, self.x_ = m, -1, beta, x
        self.xi_ = [clone(x) for _ in range(m)]
        self.fi_ = [clone(x) for _ in range(m)]
        self.tmp_ = clone(x) if m > 1 else None
        self.mat_ = np.zeros((m, m))     # mat_[j*m + i] of the reference = mat_[i, j] here
        self.rhs_ = np.zeros(m)
        self.theta_ = np.zeros(m)
        self.messages = []

    def restart(self):
        self.mm_ = -1

    def _solve(self):
        """src/AndersonMix.cc:113-236: drop history until the scaled matrix is
        well conditioned and every theta lies in [MIN_THETA, MAX_THETA]."""
        while True:
            redo = False
            while self.mm_ > 1:                                     # :134-165
                mm = self.mm_
                a = np.tril(self.mat_[:mm, :mm])
                a = a + np.tril(a, -1).T
                d = 1.0 / np.sqrt(np.diag(a))
                w = np.linalg.eigvalsh(a * d[:, None] * d[None, :])  # DSYEV on D^-1/2 A D^-1/2
                det = float(np.prod(w))
                if det < MIN_DET_MAT:
                    self.messages.append("Det. Anderson matrix=%g, set m=%d" % (det, mm))
                    self.mm_ -= 1
                else:
                    break
            mm = self.mm_
            a = np.tril(self.mat_[:mm, :mm])
            a = a + np.tril(a, -1).T
            c = np.linalg.cholesky(a)                                # DPOTRF / DPOTRS :178-181
            y = np.linalg.solve(c, self.rhs_[:mm])
            self.theta_[:mm] = np.linalg.solve(c.T, y)
            for j in range(mm):                                      # :190-233
                t = self.theta_[j]
                if t > MAX_THETA:
                    if self.mm_ > 1:
                        self.mm_ -= 1
                        redo = True
                        break
                    self.theta_[j] = -0.5 if t > 1.0 else 0.0
                elif t < MIN_THETA:
                    if self.mm_ > 1:
                        self.mm_ -= 1
                        redo = True
                        break
                    self.theta_[j] = MIN_THETA
            if not redo:
                return

    def update(self, f, work):
        """x <- Anderson-extrapolated trial solution given the residual f
        (which is replaced by the mixed residual); `work` is scratch."""
        m = self.m_
        if self.mm_ < m:
            self.mm_ += 1
        if self.mm_ > 0:                                             # :86-111
            for i in range(self.mm_):
                work.assign(f)
                work -= self.fi_[i]
                self.mat_[i, i] = work.dotProduct(work)
                self.rhs_[i] = work.dotProduct(f)
                for j in range(i):
                    self.tmp_.assign(f)
                    self.tmp_ -= self.fi_[j]
                    self.mat_[i, j] = work.dotProduct(self.tmp_)
            self._solve()
        mm = self.mm_
        if m > 0:
            for cur, hist in ((self.x_, self.xi_), (f, self.fi_)):   # :247-295
                work.assign(cur)
                if mm > 0:
                    cur.scal(self._factor(mm))
                for j in range(mm):
                    cur.axpy(float(self.theta_[j]), hist[j])
                hist.insert(0, hist.pop())
                hist[0].assign(work)
        self.x_.axpy(self.beta_ if mm > 0 else 1.0, f)               # :300-303

    def _factor(self, mm):
        factor = 1.0
        for j in range(mm):
            factor -= float(self.theta_[j])
        return factor
