// train_mnist_cnn.cpp — port of the reference's examples/train_mnist_cnn.rs against the B200 backend.
// Same model (5x Conv2dReLU 3x3 + 2x MaxPool2d + AdaptiveAvgPool2d::global + Flatten + 3x Linear, :35-100),
// Adam(lr 0.01, wd 1e-4) (:108-109), lr *= 0.8 every 5 epochs (:132-137), images reshaped [B,784] ->
// [B,1,28,28] per batch (:161-162), throughput printed per epoch (:257-258).
//
//   build/train_mnist_cnn [--data-dir DIR] [--synthetic N] [--epochs E] [--batch B] [--full-adjoint]
// By default the conv layers keep the reference's cut autograd chain (conv weights receive no
// gradient); --full-adjoint restores dW / dX.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include "taper.hpp"

using namespace taper;
using taper::data::DataLoader;
using taper::data::MNISTDataset;
using taper::loss::accuracy;
using taper::loss::cross_entropy_loss;
using namespace taper::nn;
using taper::optim::Adam;
using taper::train::Trainer;

static std::string repeat(const char* s, int n) { std::string r; while (n-- > 0) r += s; return r; }

static std::shared_ptr<Module> conv_relu(size_t cin, size_t cout, uint64_t seed) {
    return std::make_shared<Conv2dReLU>(cin, cout, Pair{3, 3}, Pair{1, 1}, Pair{1, 1}, std::nullopt, std::nullopt, true, seed);
}

int main(int argc, char** argv) {
    std::string data_dir = "./data/mnist";
    size_t synthetic = 0, epochs = 50, batch_size = 256;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--data-dir") && i + 1 < argc) data_dir = argv[++i];
        else if (!strcmp(argv[i], "--synthetic") && i + 1 < argc) synthetic = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--epochs") && i + 1 < argc) epochs = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch_size = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--full-adjoint")) Config::conv_full_adjoint() = true;
        else { fprintf(stderr, "usage: %s [--data-dir DIR] [--synthetic N] [--epochs E] [--batch B] [--full-adjoint]\n", argv[0]); return 2; }
    }
    try {
        printf("CNN MNIST Training with Performance Optimization\n\nLoading MNIST dataset...\n");
        MNISTDataset train_dataset = synthetic ? MNISTDataset::synthetic(synthetic, 1) : MNISTDataset(true, data_dir);
        MNISTDataset test_dataset = synthetic ? MNISTDataset::synthetic(synthetic / 6 + 1, 2) : MNISTDataset(false, data_dir);
        printf("Training set: %zu samples\nTest set: %zu samples\n\n", train_dataset.len(), test_dataset.len());
        DataLoader train_loader(std::move(train_dataset), batch_size, true);
        DataLoader test_loader(std::move(test_dataset), batch_size, false);

        printf("Building optimized CNN model...\n");
        auto model = std::make_shared<Sequential>(std::vector<std::shared_ptr<Module>>{
            conv_relu(1, 32, 21),                                                       // 28x28x1  -> 28x28x32
            conv_relu(32, 32, 22),                                                      // 28x28x32 -> 28x28x32
            std::make_shared<MaxPool2d>(Pair{2, 2}, Pair{2, 2}, std::nullopt),          //          -> 14x14x32
            conv_relu(32, 64, 23),
            conv_relu(64, 64, 24),
            std::make_shared<MaxPool2d>(Pair{2, 2}, Pair{2, 2}, std::nullopt),          //          -> 7x7x64
            conv_relu(64, 128, 25),
            std::make_shared<AdaptiveAvgPool2d>(AdaptiveAvgPool2d::global()),           //          -> 1x1x128
            std::make_shared<Flatten>(1),
            std::make_shared<Linear>(128, 128, true, 26),
            std::make_shared<ReLU>(),
            std::make_shared<Linear>(128, 64, true, 27),
            std::make_shared<ReLU>(),
            std::make_shared<Linear>(64, 10, true, 28),
        });
        auto params = model->parameters();
        size_t total_params = 0;
        for (auto& p : params) total_params += p.data().size();
        printf("Total parameters: %zu\n", total_params);

        float learning_rate = 0.01f;
        auto optimizer = std::make_shared<Adam>(params, learning_rate, std::nullopt, std::nullopt, 0.0001f);
        Trainer trainer(model, optimizer, nullptr);
        const size_t log_interval = 50;
        printf("\nTraining Configuration:\n   Batch size: %zu\n   Learning rate: %g\n   Epochs: %zu\n\n%s\n\n", batch_size,
               learning_rate, epochs, repeat("=", 60).c_str());

        auto total_start = std::chrono::steady_clock::now();
        std::vector<float> himg, hlab;
        size_t b = 0;
        for (size_t epoch = 1; epoch <= epochs; ++epoch) {
            auto epoch_start = std::chrono::steady_clock::now();
            printf("Epoch %zu/%zu\n", epoch, epochs);
            if (epoch % 5 == 0 && epoch >= 5) {
                learning_rate *= 0.8f;
                printf("   Reducing learning rate to %.6f\n", learning_rate);
                trainer.optimizer->set_lr(learning_rate);
            }
            for (auto& p : trainer.model->parameters()) p.zero_grad();

            float train_loss = 0.0f;
            size_t train_correct = 0, train_total = 0, batch_idx = 0;
            std::vector<double> batch_times;
            train_loader.reset();
            size_t num_batches = train_loader.num_batches();
            while (train_loader.next(himg, hlab, b)) {
                auto batch_start = std::chrono::steady_clock::now();
                Tape::reset();
                Tensor images = Tensor::from_host(himg.data(), {b, 784});
                Tensor labels = Tensor::from_host(hlab.data(), {b});
                Tensor images_4d = images.reshape({b, 1, 28, 28});
                Tensor logits = trainer.model->forward(images_4d);
                Tensor loss = cross_entropy_loss(logits, labels);
                float batch_acc = accuracy(logits, labels);
                train_correct += (size_t)(batch_acc * (float)b);
                train_total += b;
                loss.backward();
                trainer.optimizer->step();
                trainer.optimizer->zero_grad();
                train_loss += loss.data()[0];
                batch_times.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - batch_start).count());
                if ((batch_idx + 1) % log_interval == 0 || batch_idx == num_batches - 1) {
                    double avg = std::accumulate(batch_times.begin(), batch_times.end(), 0.0) / (double)batch_times.size();
                    printf("\r   Batch [%zu/%zu] Loss: %.4f, Acc: %.2f%%, Avg Batch Time: %.3fms", batch_idx + 1, num_batches,
                           loss.data()[0], 100.0f * (float)train_correct / (float)train_total, avg);
                    fflush(stdout);
                }
                ++batch_idx;
            }
            float avg_train_loss = train_loss / (float)num_batches;
            float train_accuracy = (float)train_correct / (float)train_total;
            printf("\n   Evaluating...");
            fflush(stdout);

            auto val_start = std::chrono::steady_clock::now();
            float val_loss = 0.0f;
            size_t val_correct = 0, val_total = 0;
            test_loader.reset();
            size_t num_val_batches = test_loader.num_batches();
            while (test_loader.next(himg, hlab, b)) {
                Tape::reset();
                Tensor images = Tensor::from_host(himg.data(), {b, 784});
                Tensor labels = Tensor::from_host(hlab.data(), {b});
                Tensor logits = trainer.model->forward(images.reshape({b, 1, 28, 28}));
                Tensor loss = cross_entropy_loss(logits, labels);
                float batch_acc = accuracy(logits, labels);
                val_correct += (size_t)(batch_acc * (float)b);
                val_total += b;
                val_loss += loss.data()[0];
            }
            float avg_val_loss = val_loss / (float)num_val_batches;
            float val_accuracy = (float)val_correct / (float)val_total;
            auto now = std::chrono::steady_clock::now();
            float epoch_time = std::chrono::duration<float>(now - epoch_start).count();
            double val_ms = std::chrono::duration<double, std::milli>(now - val_start).count();
            printf("\rEpoch %zu complete:\n", epoch);
            printf("   Train Loss: %.4f | Train Acc: %.2f%%\n", avg_train_loss, train_accuracy * 100.0f);
            printf("   Val Loss: %.4f   | Val Acc: %.2f%%\n", avg_val_loss, val_accuracy * 100.0f);
            printf("   Time: %.2fs (Val: %.0fms)\n", epoch_time, val_ms);
            printf("   Throughput: %.0f samples/sec\n\n", (float)train_total / epoch_time);
            if (val_accuracy > 0.995f) {
                printf("Reached %.2f%% validation accuracy! Stopping early.\n", val_accuracy * 100.0f);
                break;
            }
        }
        float total_time = std::chrono::duration<float>(std::chrono::steady_clock::now() - total_start).count();
        printf("\n%s\nTraining Complete! Total time: %.2fs\n\nTesting CNN on sample images:\n", repeat("=", 60).c_str(), total_time);
        test_loader.reset();
        if (test_loader.next(himg, hlab, b)) {
            Tensor images = Tensor::from_host(himg.data(), {b, 784});
            Tensor predictions = trainer.model->forward(images.reshape({b, 1, 28, 28}));
            Tensor pred_classes = predictions.argmax(1);
            for (size_t i = 0; i < std::min<size_t>(10, b); ++i) {
                int predicted = (int)pred_classes.data()[i], actual = (int)hlab[i];
                printf("Sample %zu: Predicted=%d, Actual=%d %s\n", i + 1, predicted, actual, predicted == actual ? "Correct" : "False");
            }
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
