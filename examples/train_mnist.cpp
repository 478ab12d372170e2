// train_mnist.cpp — port of the reference's examples/train_mnist.rs against the B200 backend.
// Same model (MLP 784-128-64-10, :32-40), same optimizer (Adam lr 1e-3, wd 1e-4, :50-51), same manual
// epoch loop (:69-197): Tape::reset, forward, cross_entropy_loss, accuracy, backward, step, zero_grad,
// loss.data()[0] read back every batch.  Every tensor op below launches a hand-written sm_100a kernel
// through the C ABI; nothing runs on the host except the loop itself.
//
//   build/train_mnist [--data-dir DIR] [--synthetic N] [--epochs E] [--batch B]
// The reference downloads MNIST when the IDX files are missing (src/data/mnist.rs:60-181); there is no
// network here, so --synthetic N substitutes N MNIST-shaped random samples (and N/6 for the test set).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "taper.hpp"

using namespace taper;
using taper::data::DataLoader;
using taper::data::MNISTDataset;
using taper::loss::accuracy;
using taper::loss::cross_entropy_loss;
using taper::nn::Linear;
using taper::nn::Module;
using taper::nn::ReLU;
using taper::nn::Sequential;
using taper::optim::Adam;
using taper::train::Trainer;

static std::string repeat(const char* s, int n) { std::string r; while (n-- > 0) r += s; return r; }

int main(int argc, char** argv) {
    std::string data_dir = "./data/mnist";
    size_t synthetic = 0, epochs = 10, batch_size = 256;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--data-dir") && i + 1 < argc) data_dir = argv[++i];
        else if (!strcmp(argv[i], "--synthetic") && i + 1 < argc) synthetic = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--epochs") && i + 1 < argc) epochs = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch_size = strtoull(argv[++i], nullptr, 10);
        else { fprintf(stderr, "usage: %s [--data-dir DIR] [--synthetic N] [--epochs E] [--batch B]\n", argv[0]); return 2; }
    }
    try {
        printf("MNIST Neural Network Training\n\n");
        printf("Loading MNIST dataset...\n");
        MNISTDataset train_dataset = synthetic ? MNISTDataset::synthetic(synthetic, 1) : MNISTDataset(true, data_dir);
        MNISTDataset test_dataset = synthetic ? MNISTDataset::synthetic(synthetic / 6 + 1, 2) : MNISTDataset(false, data_dir);
        printf("Training set: %zu samples\n", train_dataset.len());
        printf("Test set: %zu samples\n\n", test_dataset.len());

        DataLoader train_loader(std::move(train_dataset), batch_size, true);
        DataLoader test_loader(std::move(test_dataset), batch_size, false);

        printf("Building model...\n");
        auto model = std::make_shared<Sequential>(std::vector<std::shared_ptr<Module>>{
            std::make_shared<Linear>(784, 128, true, 11),
            std::make_shared<ReLU>(),
            std::make_shared<Linear>(128, 64, true, 12),
            std::make_shared<ReLU>(),
            std::make_shared<Linear>(64, 10, true, 13),
        });
        auto params = model->parameters();
        size_t total = 0;
        for (auto& p : params) total += p.data().size();
        printf("Total parameters: %zu\n", total);

        float learning_rate = 0.001f;
        auto optimizer = std::make_shared<Adam>(params, learning_rate, std::nullopt, std::nullopt, 0.0001f);
        Trainer trainer(model, optimizer, nullptr);

        printf("\nTraining Configuration:\n   Batch size: %zu\n   Learning rate: %g\n   Epochs: %zu\n\n%s\n\n", batch_size,
               learning_rate, epochs, repeat("=", 60).c_str());

        std::vector<float> himg, hlab;
        size_t b = 0;
        for (size_t epoch = 1; epoch <= epochs; ++epoch) {
            auto epoch_start = std::chrono::steady_clock::now();
            printf("Epoch %zu/%zu\n", epoch, epochs);
            for (auto& p : trainer.model->parameters()) p.zero_grad();

            float train_loss = 0.0f;
            size_t train_correct = 0, train_total = 0, batch_idx = 0;
            train_loader.reset();
            size_t num_batches = train_loader.num_batches();
            while (train_loader.next(himg, hlab, b)) {
                Tape::reset();
                Tensor images = Tensor::from_host(himg.data(), {b, 784});
                Tensor labels = Tensor::from_host(hlab.data(), {b});
                Tensor logits = trainer.model->forward(images);
                if (logits.shape()[1] != 10) throw std::runtime_error("Output should have 10 classes");
                Tensor loss = cross_entropy_loss(logits, labels);
                float batch_acc = accuracy(logits, labels);
                train_correct += (size_t)(batch_acc * (float)b);
                train_total += b;
                loss.backward();
                trainer.optimizer->step();
                trainer.optimizer->zero_grad();
                train_loss += loss.data()[0];
                if ((batch_idx + 1) % 100 == 0 || batch_idx == num_batches - 1) {
                    printf("\r   Batch [%zu/%zu] Loss: %.4f, Acc: %.2f%%", batch_idx + 1, num_batches, loss.data()[0],
                           100.0f * (float)train_correct / (float)train_total);
                    fflush(stdout);
                }
                ++batch_idx;
            }
            float avg_train_loss = train_loss / (float)num_batches;
            float train_accuracy = (float)train_correct / (float)train_total;
            printf("\n   Evaluating...");
            fflush(stdout);

            float val_loss = 0.0f;
            size_t val_correct = 0, val_total = 0;
            test_loader.reset();
            size_t num_val_batches = test_loader.num_batches();
            while (test_loader.next(himg, hlab, b)) {
                Tape::reset();
                Tensor images = Tensor::from_host(himg.data(), {b, 784});
                Tensor labels = Tensor::from_host(hlab.data(), {b});
                Tensor logits = trainer.model->forward(images);
                Tensor loss = cross_entropy_loss(logits, labels);
                float batch_acc = accuracy(logits, labels);
                val_correct += (size_t)(batch_acc * (float)b);
                val_total += b;
                val_loss += loss.data()[0];
            }
            float avg_val_loss = val_loss / (float)num_val_batches;
            float val_accuracy = (float)val_correct / (float)val_total;
            float epoch_time = std::chrono::duration<float>(std::chrono::steady_clock::now() - epoch_start).count();
            printf("\rEpoch %zu complete:\n", epoch);
            printf("   Train Loss: %.4f | Train Acc: %.2f%%\n", avg_train_loss, train_accuracy * 100.0f);
            printf("   Val Loss: %.4f   | Val Acc: %.2f%%\n", avg_val_loss, val_accuracy * 100.0f);
            printf("   Time: %.2fs  (%.0f train samples/s incl. eval)\n\n", epoch_time, (float)train_total / epoch_time);
            if (val_accuracy > 0.98f) {
                printf("Reached %.2f%% validation accuracy! Stopping early.\n", val_accuracy * 100.0f);
                break;
            }
        }
        printf("\n%s\nTraining Complete!\n\nTesting on sample images:\n", repeat("=", 60).c_str());
        test_loader.reset();
        if (test_loader.next(himg, hlab, b)) {
            Tensor images = Tensor::from_host(himg.data(), {b, 784});
            Tensor predictions = trainer.model->forward(images);
            Tensor pred_classes = predictions.argmax(1);
            for (size_t i = 0; i < std::min<size_t>(5, b); ++i) {
                int predicted = (int)pred_classes.data()[i], actual = (int)hlab[i];
                printf("Sample %zu: Predicted=%d, Actual=%d %s\n", i + 1, predicted, actual, predicted == actual ? "Correct" : "Wrong");
            }
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
