// xor.cpp — port of the reference's src/main.rs (XOR demo): 2-4-1 MLP with Sigmoid activations, BCE loss, SGD(lr 0.10),
// 50 000 iterations through the eager tape (every op one kernel launch through the C ABI), loss printed every 1000.
#include <cstdio>
#include <cstdlib>
#include "taper.hpp"

using namespace taper;

int main(int argc, char** argv) {
    size_t epochs = argc > 1 ? strtoull(argv[1], nullptr, 10) : 50000;
    try {
        printf("XOR Neural Network Training\n\n");
        const std::vector<float> x_data = {0, 0, 0, 1, 1, 0, 1, 1};
        const std::vector<float> y_data = {0, 1, 1, 0};
        auto model = std::make_shared<nn::Sequential>(std::vector<std::shared_ptr<nn::Module>>{
            std::make_shared<nn::Linear>(2, 4, true, 3),
            std::make_shared<nn::Sigmoid>(),
            std::make_shared<nn::Linear>(4, 1, true, 4),
            std::make_shared<nn::Sigmoid>(),
        });
        optim::SGD opt(model->parameters(), 0.10f, std::nullopt);
        for (size_t epoch = 0; epoch < epochs; ++epoch) {
            Tape::reset();
            Tensor x = Tensor::create(x_data, {4, 2});
            Tensor y = Tensor::create(y_data, {4, 1});
            Tensor yhat = model->forward(x);
            Tensor loss = loss::bce_loss(yhat, y);
            loss.backward();
            opt.step();
            opt.zero_grad();
            if (epoch % 1000 == 0) printf("iteration %4zu: Loss = %.4f\n", epoch, loss.data()[0]);
        }
        Tape::reset();
        Tensor yhat = model->forward(Tensor::create(x_data, {4, 2}));
        const std::vector<float>& p = yhat.data();
        printf("\n[0,0]->%.3f\n[0,1]->%.3f\n[1,0]->%.3f\n[1,1]->%.3f\n", p[0], p[1], p[2], p[3]);
        bool ok = p[0] < 0.5f && p[1] > 0.5f && p[2] > 0.5f && p[3] < 0.5f;
        printf("%s\n", ok ? "learned XOR" : "not yet");
        return ok ? 0 : 3;
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
