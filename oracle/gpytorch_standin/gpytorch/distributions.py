class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.mean, self.covariance_matrix = mean, covariance_matrix
